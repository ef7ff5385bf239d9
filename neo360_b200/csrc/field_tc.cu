// tcgen05 tensor-core path (placeholder until the kernel lands)
#include "common.cuh"
namespace neo {
int tc_scene_create(NeoScene* sc, const NeoMLPParams mlps[4], cudaStream_t s) { set_error("NEO_PREC_TC not built yet"); return NEO_ERR_UNSUPPORTED; }
void tc_scene_free(NeoScene* sc) {}
int launch_field_tc(const NeoScene* sc, const NeoRays* rays, const float* far, const float* t, int N, int mlp_index, float* rgb, float* sigma, cudaStream_t s) { set_error("NEO_PREC_TC not built yet"); return NEO_ERR_UNSUPPORTED; }
}
