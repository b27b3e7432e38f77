// TEST INFRASTRUCTURE ONLY -- C-ABI shim over the reference's points_in_boxes launcher
// (mmdet3d/ops/roiaware_pool3d/src/points_in_boxes_cuda.cu:127-149), compiled UNMODIFIED from where it lies together with this
// file into oracle/_ref/libref_pib.so.  That file includes torch headers for its at::Tensor wrappers, so this library is
// built separately from libref_ops.so and only where the torch headers / libraries of the image are found.
void points_in_boxes_batch_launcher(int batch_size, int boxes_num, int pts_num, const float *boxes, const float *pts,
                                    int *box_idx_of_points);
extern "C" void ref_points_in_boxes_batch(int batch_size, int boxes_num, int pts_num, const float *boxes, const float *pts, int *out) {
  points_in_boxes_batch_launcher(batch_size, boxes_num, pts_num, boxes, pts, out);
}
