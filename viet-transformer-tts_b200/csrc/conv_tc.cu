// conv_tc.cu -- tcgen05 path (placeholder until the implicit-GEMM kernels land).
#include "generator.cuh"
namespace vtts {
int tc_pack_layer(VttsGen *, int, cudaStream_t) { return VTTS_OK; }
int tc_workspace_bytes(const VttsGen *, int, int, size_t *) {
    return set_error(VTTS_E_UNSUPPORTED, "bf16 tcgen05 path not built");
}
int tc_forward(VttsGen *, const float *, const float *, float *, int, int, void *, size_t, int, float *, cudaStream_t) {
    return set_error(VTTS_E_UNSUPPORTED, "bf16 tcgen05 path not built");
}
void tc_destroy(VttsGen *) {}
int tc_supported(const VttsGen *, char *, size_t) { return 0; }
}  // namespace vtts
extern "C" int vtts_dbg_umma_gemm(const void *, const void *, float *, int, int, int, int, int, int, vtts_stream_t) {
    return vtts::set_error(VTTS_E_UNSUPPORTED, "umma probe not built");
}
