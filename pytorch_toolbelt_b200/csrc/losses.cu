// Segmentation-loss kernels (placeholder translation unit: real kernels land in the next milestone).
#include "common.cuh"
using namespace tb;
extern "C" {
size_t tb_loss_workspace_bytes(int64_t, int, int64_t) { return 0; }
size_t tb_lovasz_workspace_bytes(int64_t, int) { return 0; }
}
