// libstage_b200.so — C-ABI entry points (include/stage_b200.h).  One translation unit: the kernels live in the
// .cuh files next to this one.
#include "common.cuh"
#include "umma.cuh"
#include "probe.cuh"

using namespace stg;

extern "C" {

int stage_abi_version(void) { return STAGE_ABI_VERSION; }
const char* stage_last_error(void) { return err_buf(); }

int stage_check_device_error(void* stream) {
  unsigned int h = 0, zero = 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return fail(STAGE_E_LAUNCH, "stream sync: %s", cudaGetErrorString(e));
  e = cudaMemcpyFromSymbol(&h, g_dev_error, sizeof(h));
  if (e != cudaSuccess) return fail(STAGE_E_LAUNCH, "read error flag: %s", cudaGetErrorString(e));
  if (h != 0) {
    cudaMemcpyToSymbol(g_dev_error, &zero, sizeof(zero));
    return fail(STAGE_E_DEVICE, "device error flag 0x%x (bounded mbarrier wait timed out)", h);
  }
  return STAGE_OK;
}

int stage_umma_probe(const StageProbeArgs* a, void* stream) {
  STG_REQUIRE(a && a->a_img && a->b_img && a->d_out, STAGE_E_BADARG, "probe: null pointer");
  STG_REQUIRE(a->n % 16 == 0 && a->n >= 16 && a->n <= 256, STAGE_E_UNSUPPORTED, "probe: n=%u", a->n);
  ProbeArgs p;
  p.a_img = a->a_img; p.b_img = a->b_img; p.a_bytes = a->a_bytes; p.b_bytes = a->b_bytes;
  p.a_lbo = a->a_lbo; p.a_sbo = a->a_sbo; p.a_step = a->a_step;
  p.b_lbo = a->b_lbo; p.b_sbo = a->b_sbo; p.b_step = a->b_step;
  p.idesc = make_idesc_f16(a->n, a->a_mn, a->b_mn);
  p.k_steps = a->k_steps; p.n = a->n; p.use_bulk = a->use_bulk; p.d_out = a->d_out;
  size_t smem = ((a->a_bytes + 1023u) & ~1023u) + a->b_bytes + 1024;
  STG_REQUIRE(smem <= 200 * 1024, STAGE_E_UNSUPPORTED, "probe: smem %zu", smem);
  int rc = set_smem(umma_probe_kernel, smem, "probe");
  if (rc) return rc;
  umma_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("umma_probe");
}

// ---- not built yet (filled in as the kernels land) ----
#ifndef STG_HAVE_PACK
size_t stage_packed_weight_bytes(int, int, int) { return 0; }
int stage_pack_weight(const float*, int, int, int, void*, void*) { return fail(STAGE_E_UNSUPPORTED, "not built"); }
#endif
#ifndef STG_HAVE_LINEAR
int stage_linear_fwd(const StageLinearArgs*, void*) { return fail(STAGE_E_UNSUPPORTED, "not built"); }
#endif
#ifndef STG_HAVE_ENCODER
int stage_encoder_block_fwd(const StageEncoderArgs*, void*) { return fail(STAGE_E_UNSUPPORTED, "not built"); }
#endif
#ifndef STG_HAVE_CQA
int stage_cqa_fused_fwd(const StageCqaArgs*, void*) { return fail(STAGE_E_UNSUPPORTED, "not built"); }
#endif
#ifndef STG_HAVE_HEAD
int stage_span_head_fwd(const StageHeadArgs*, void*) { return fail(STAGE_E_UNSUPPORTED, "not built"); }
int stage_find_max_span(const float*, const float*, int, int, int32_t*, float*, void*) { return fail(STAGE_E_UNSUPPORTED, "not built"); }
int stage_mask_logits(const float*, const float*, int64_t, int, float*, void*) { return fail(STAGE_E_UNSUPPORTED, "not built"); }
#endif

}  // extern "C"
