// The rel-1e-4 variant of the projected-map tile kernel: field_bin.cu compiled a second time with every operand carried as
// an fp16 (hi, lo) pair and every product as three tensor-core products (namespace sd::tbx, launch_field_bin_x3).
//   BTSNet.forward      models/bts.py:476-595 at the reference's own evaluation precision (fp32,
//                       configs/evaluate_semantic_kitti_360.yaml:15; resnetfc.py:163,199)
#define SD_TB_X3 1
#include "field_bin.cu"
