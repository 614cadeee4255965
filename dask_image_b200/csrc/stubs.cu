// stubs.cu -- kernel families not built yet report "not eligible" (-1) so the
// API falls through to the generic kernels.
#include "common.cuh"
namespace b200 {
int launch_corrnd_tiled(const CorrNdArgs &) { return -1; }
}  // namespace b200
