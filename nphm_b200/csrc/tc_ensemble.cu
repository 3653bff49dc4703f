// placeholder until the tcgen05 kernel lands
#include "engine.cuh"
namespace nphm {
bool tc_ensemble_supported(const nphm_ensemble *) { return false; }
int tc_ensemble_pack(nphm_ensemble *, cudaStream_t) { return NPHM_OK; }
int tc_ensemble_launch(nphm_ensemble *, const SimtQuery &, cudaStream_t) { set_error("tcgen05 kernel not built"); return NPHM_ERR_UNSUPPORTED; }
}
