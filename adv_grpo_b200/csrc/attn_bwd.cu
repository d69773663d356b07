// Flash-attention backward (placeholder until the tcgen05 kernel lands in this file).
#include "common.cuh"

using namespace advgrpo;

extern "C" {

size_t advgrpo_attn_bwd_workspace_bytes(int64_t B, int64_t S, int64_t H, int64_t D) {
  (void)D;
  return (size_t)B * H * S * sizeof(float) + 256;
}

int advgrpo_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                     void* dqkv, int64_t B, int64_t S, int64_t H, int64_t D, float scale,
                     int causal, void* workspace, size_t workspace_bytes,
                     advgrpo_stream_t stream) {
  (void)qkv; (void)out; (void)dout; (void)lse; (void)dqkv; (void)B; (void)S; (void)H; (void)D;
  (void)scale; (void)causal; (void)workspace; (void)workspace_bytes; (void)stream;
  return set_error(ADVGRPO_ERR_UNSUPPORTED, "attn_bwd: not implemented yet");
}

}  // extern "C"
