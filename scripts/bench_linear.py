"""FFMA vs the three tensor-core generations of cn_linear across shapes (decides the dispatch thresholds in kernels.cn_linear).
Prints one JSON line per shape: time, TFLOP/s and achieved fraction of the HBM roofline (algorithmic bytes: X in + Y out + W)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import pcreid_b200.kernels as K  # noqa: E402

dev = "cuda"
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM = 6545.6


def timeit(fn, iters=5):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


SHAPES = [(2048, 256, 32, 32), (2048, 256, 64, 64), (2048, 256, 64, 128), (2048, 256, 64, 192), (2048, 256, 128, 128), (2048, 256, 128, 256),
          (2048, 256, 256, 256), (1024, 256, 512, 1024), (1024, 256, 1024, 512), (1024, 256, 512, 128), (512, 1024, 128, 128),
          (4096, 128, 128, 128), (8192, 64, 64, 64)]
for (B, N, Kd, CO) in SHAPES:
    x = torch.randn(B, Kd, N, device=dev); w = torch.randn(Kd, CO, device=dev) / Kd ** 0.5
    out = torch.empty(B, CO, N, device=dev)
    res = {}
    with K.tensor_core_linear(False):
        res["ffma"] = timeit(lambda: K.cn_linear(x, w, act=1, out=out))
    ref = out.clone()
    K._TC_LINEAR["min_k"] = 8
    for name, cfg in (("tc1", dict(gen=1, tma=False)), ("tc2", dict(gen=2, tma=False)), ("tma", dict(tma=True)), ("tma128", dict(tma=True, tma_tile128=True)), ("x3", dict(x3=True, tma_min_k=1 << 30))):
        K._TC_LINEAR.update(cfg)
        with K.tensor_core_linear(True):
            res[name] = timeit(lambda: K.cn_linear(x, w, act=1, out=out))
        res[name + "_err"] = float((out - ref).abs().max())
        for k in cfg:
            K._TC_LINEAR.pop(k)
    fl = 2.0 * B * N * Kd * CO
    by = 4.0 * (B * N * (Kd + CO) + Kd * CO)
    print(json.dumps({"B": B, "N": N, "K": Kd, "CO": CO,
                      **{k + "_ms": round(v, 4) for k, v in res.items() if not k.endswith("_err")},
                      **{k + "_TFLOPs": round(fl / res[k] / 1e9, 1) for k in ("ffma", "tc1", "tc2", "tma", "tma128", "x3")},
                      "tma_frac_hbm": round(by / res["tma"] / 1e6 / HBM, 3), "tma_max_err": res["tma_err"], "tc2_max_err": res["tc2_err"], "x3_max_err": res["x3_err"]}), flush=True)
