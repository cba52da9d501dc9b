// tcgen05 / TMA framed contraction (placeholder until the kernel lands).
#include "common.cuh"
namespace nnab {
bool tc_supported(const FramedProblem&) { return false; }
size_t tc_workspace_bytes(int64_t, int64_t, int, int, int) { return 0; }
int launch_framed_tc(const FramedProblem&, const void*, void*, size_t, cudaStream_t) { return NNAB_EUNSUPPORTED; }
size_t tc_packed_bytes(int, int) { return 0; }
int tc_pack_basis(const float*, const float*, int, int, void*, cudaStream_t) { return NNAB_EUNSUPPORTED; }
int tc_tile_n() { return 256; }
}  // namespace nnab
