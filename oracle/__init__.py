"""Parity oracle (TEST INFRASTRUCTURE ONLY).  See reid_oracle.py / ops_oracle.c headers."""
