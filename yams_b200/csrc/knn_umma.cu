// knn_umma.cu -- tcgen05 (5th-gen tensor core) stage-1 engine. Placeholder until the UMMA kernel lands:
// reports "unsupported" so the CUDA-core engine runs.
#include "knn.cuh"
namespace yb {
bool tcgen05_supported(const Corpus*, uint32_t) { return false; }
yams_status_t stage1_tcgen05(Corpus*, const Stage1Args&, bool, cudaStream_t) { return YAMS_ERR_UNSUPPORTED; }
}  // namespace yb
