// libstage_b200.so — C-ABI entry points (include/stage_b200.h).  One translation unit: the kernels live in the
// .cuh files next to this one.
#include "common.cuh"
#include "umma.cuh"
#include "probe.cuh"
#include "pack.cuh"
#include "linear.cuh"
#include "encoder.cuh"
#include "cqa.cuh"
#include "mha.cuh"
#include "head.cuh"
#define STG_HAVE_PACK 1
#define STG_HAVE_HEAD 1
#define STG_HAVE_CQA 1
#define STG_HAVE_ENCODER 1
#define STG_HAVE_LINEAR 1

using namespace stg;

namespace {
template <int NPAD, int KC, int STAGES>
static int launch_linear(const LinearDev& d, cudaStream_t s) {
  using Cfg = LinearCfg<NPAD, KC, STAGES>;
  int rc = set_smem(linear_kernel<NPAD, KC, STAGES>, Cfg::SMEM_BYTES, "linear");
  if (rc) return rc;
  const long long tiles = (d.m + 127) / 128;
  const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
  linear_kernel<NPAD, KC, STAGES><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, s>>>(d);
  return check_launch("linear");
}

__global__ void fill_f32_kernel(float* __restrict__ p, long long n, float v) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) p[i] = v;
}

template <int KS>
static int launch_encoder(EncDev& d, cudaStream_t s) {
  constexpr size_t smem = EncCfg::smem_bytes(KS);
  const int G = 128 / d.L;
  // Packed tiling (masked max-pool output only, see encoder.cuh): worth it whenever whole sequences fill a tile worse than
  // the 128 - 2*halo exact rows of a packed tile do (L = 100: 100 rows per tile -> 120)
  const int halo = d.n_conv * (KS / 2), out_rows = 128 - 2 * halo;
  d.packed = (d.out_mode == 1 && out_rows > 0 && G * d.L < out_rows && d.n_seq * d.L < (1ll << 30)) ? 1 : 0;
  long long tiles = (d.n_seq + G - 1) / G;
  if (d.packed) tiles = (d.n_seq * d.L + out_rows - 1) / out_rows;
  const long long slots = 2ll * num_sms();  // two CTAs fit on an SM (shared memory, 256 TMEM columns each)
  const int grid = static_cast<int>(tiles < slots ? tiles : slots);
  if (d.packed) {
    int rc = set_smem(encoder_kernel<KS, true>, smem, "encoder");
    if (rc) return rc;
    const long long n = d.n_seq * 128;
    fill_f32_kernel<<<static_cast<int>((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, s>>>(d.y, n, -INFINITY);
    encoder_kernel<KS, true><<<grid, EncCfg::THREADS, smem, s>>>(d);
  } else {
    int rc = set_smem(encoder_kernel<KS, false>, smem, "encoder");
    if (rc) return rc;
    encoder_kernel<KS, false><<<grid, EncCfg::THREADS, smem, s>>>(d);
  }
  return check_launch("encoder");
}
}  // namespace

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

int stage_post_device_error(unsigned int* host_flag, void* stream) {
  STG_REQUIRE(host_flag, STAGE_E_BADARG, "post_device_error: null pointer");
  cudaError_t e = cudaMemcpyFromSymbolAsync(host_flag, g_dev_error, sizeof(unsigned int), 0, cudaMemcpyDeviceToHost,
                                            static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(STAGE_E_LAUNCH, "post error flag: %s", cudaGetErrorString(e));
  return STAGE_OK;
}

int stage_umma_probe(const StageProbeArgs* a, void* stream) {
  STG_REQUIRE(a && a->a_img && a->b_img && a->d_out, STAGE_E_BADARG, "probe: null pointer");
  STG_REQUIRE(a->n % 16 == 0 && a->n >= 16 && a->n <= 256, STAGE_E_UNSUPPORTED, "probe: n=%u", a->n);
  ProbeArgs p;
  p.a_img = a->a_img; p.b_img = a->b_img; p.a_bytes = a->a_bytes; p.b_bytes = a->b_bytes;
  p.a_lbo = a->a_lbo; p.a_sbo = a->a_sbo; p.a_step = a->a_step;
  p.b_lbo = a->b_lbo; p.b_sbo = a->b_sbo; p.b_step = a->b_step;
  p.a_tmem = (a->a_mn == 2u) ? 1u : 0u;  /* a_mn = 2: A operand from tensor memory (K-major by construction) */
  p.idesc = make_idesc_f16(a->n, p.a_tmem ? 0u : a->a_mn, a->b_mn);
  p.k_steps = a->k_steps; p.n = a->n; p.use_bulk = a->use_bulk; p.d_out = a->d_out;
  p.repeat = a->repeat > 0 ? a->repeat : 1u; p.cycles_out = a->cycles_out;
  size_t smem = ((a->a_bytes + 1023u) & ~1023u) + a->b_bytes + 1024;
  STG_REQUIRE(smem <= 200 * 1024, STAGE_E_UNSUPPORTED, "probe: smem %zu", smem);
  int rc = set_smem(umma_probe_kernel, smem, "probe");
  if (rc) return rc;
  umma_probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("umma_probe");
}

size_t stage_packed_weight_bytes(int n, int k, int kc) {
  if (n <= 0 || n > 304 || k <= 0 || (kc != 32 && kc != 64)) return 0;
  PackGeom g = pack_geom(n, k, kc);
  return static_cast<size_t>(g.n_chunks) * g.chunk_bytes;
}

int stage_pack_weight(const float* w, int n, int k, int kc, void* packed, void* stream) {
  STG_REQUIRE(w && packed, STAGE_E_BADARG, "pack_weight: null pointer");
  STG_REQUIRE(n > 0 && n <= 304 && k > 0 && (kc == 32 || kc == 64), STAGE_E_UNSUPPORTED, "pack_weight: n=%d k=%d kc=%d", n, k, kc);
  STG_REQUIRE(aligned16(packed), STAGE_E_BADARG, "pack_weight: packed buffer not 16-byte aligned");
  PackGeom g = pack_geom(n, k, kc);
  const int total = (g.kpad / 8) * g.npad;
  pack_weight_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, k, 0, nullptr, g, static_cast<uint8_t*>(packed));
  return check_launch("pack_weight");
}

int stage_pack_weight_ln(const float* w, int n, int k, int kc, const float* ln_g, const float* ln_b, const float* bias,
                         void* packed, float* cs, float* cb, void* stream) {
  STG_REQUIRE(w && ln_g && ln_b && bias && packed && cs && cb, STAGE_E_BADARG, "pack_weight_ln: null pointer");
  STG_REQUIRE(n > 0 && n <= 304 && k > 0 && (kc == 32 || kc == 64), STAGE_E_UNSUPPORTED, "pack_weight_ln: n=%d k=%d kc=%d", n, k, kc);
  STG_REQUIRE(aligned16(packed), STAGE_E_BADARG, "pack_weight_ln: packed buffer not 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PackGeom g = pack_geom(n, k, kc);
  const int total = (g.kpad / 8) * g.npad;
  pack_weight_kernel<<<(total + 255) / 256, 256, 0, s>>>(w, k, 0, ln_g, g, static_cast<uint8_t*>(packed));
  fold_layernorm_kernel<<<(n * 32 + 255) / 256, 256, 0, s>>>(w, n, k, ln_g, ln_b, bias, cs, cb);
  return check_launch("pack_weight_ln");
}

int stage_linear_fwd(const StageLinearArgs* a, void* stream) {
  STG_REQUIRE(a && a->x && a->w_packed && a->y, STAGE_E_BADARG, "linear: null pointer");
  STG_REQUIRE(a->m > 0 && a->k > 0 && a->n > 0, STAGE_E_BADARG, "linear: m=%lld k=%d n=%d", (long long)a->m, a->k, a->n);
  STG_REQUIRE(a->k % 4 == 0, STAGE_E_UNSUPPORTED, "linear: k=%d must be a multiple of 4", a->k);
  STG_REQUIRE(a->n <= 304, STAGE_E_UNSUPPORTED, "linear: n=%d > 304", a->n);
  STG_REQUIRE(aligned16(a->x) && aligned16(a->y) && aligned16(a->w_packed), STAGE_E_BADARG, "linear: pointers must be 16-byte aligned");
  STG_REQUIRE((a->ln_in_cs == nullptr) == (a->ln_in_cb == nullptr) && (a->ln_out_g == nullptr) == (a->ln_out_b == nullptr),
              STAGE_E_BADARG, "linear: LayerNorm vectors must come in pairs");
  STG_REQUIRE(a->ln_in_cs != nullptr || a->bias != nullptr, STAGE_E_BADARG, "linear: null bias");
  const int want_kc = a->n <= 128 ? 64 : 32;
  STG_REQUIRE(a->kc == want_kc, STAGE_E_UNSUPPORTED, "linear: weight for n=%d must be packed with kc=%d", a->n, want_kc);
  PackGeom g = pack_geom(a->n, a->k, a->kc);
  LinearDev d;
  d.x = a->x; d.w = static_cast<const uint8_t*>(a->w_packed); d.bias = a->bias; d.cs = a->ln_in_cs; d.cb = a->ln_in_cb;
  d.ln_out_g = a->ln_out_g; d.ln_out_b = a->ln_out_b;
  d.y = a->y; d.m = a->m; d.k = a->k; d.n = a->n; d.n_chunks = g.n_chunks; d.l2norm_in = a->l2norm_in; d.relu = a->relu;
  d.out_blocked = 0; d.rows_per_block = 1; d.timeline = a->timeline;
  d.ldy = a->ldy > 0 ? a->ldy : a->n;
  d.x_rowmax = a->x_rowmax;
  STG_REQUIRE(d.ldy >= a->n, STAGE_E_BADARG, "linear: ldy=%lld < n=%d", (long long)d.ldy, a->n);
  STG_REQUIRE(a->x_rowmax == nullptr || (a->ln_in_cs == nullptr && !a->l2norm_in), STAGE_E_UNSUPPORTED, "linear: x_rowmax cannot be combined with LN_in / l2norm_in");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (a->n <= 128) return launch_linear<128, 64, 2>(d, s);
  return launch_linear<304, 32, 3>(d, s);
}

int stage_blocked_stride(int L) { return L <= 0 ? 0 : L + ((9 - L % 8) % 8); }

int stage_encoder_block_fwd(const StageEncoderArgs* a, void* stream) {
  STG_REQUIRE(a && a->x && a->y, STAGE_E_BADARG, "encoder: null pointer");
  STG_REQUIRE((a->flags & STAGE_ENC_NO_PE) || a->pe, STAGE_E_BADARG, "encoder: null position-encoding table");
  STG_REQUIRE((a->flags & STAGE_ENC_NO_FINAL_LN) || (a->final_ln_g && a->final_ln_b), STAGE_E_BADARG, "encoder: null final LayerNorm");
  STG_REQUIRE(a->n_seq > 0 && a->L > 0, STAGE_E_BADARG, "encoder: n_seq=%lld L=%d", (long long)a->n_seq, a->L);
  STG_REQUIRE(a->L <= STAGE_MAX_SEQ, STAGE_E_UNSUPPORTED, "encoder: L=%d > %d", a->L, STAGE_MAX_SEQ);
  STG_REQUIRE(a->n_conv >= 1 && a->n_conv <= STAGE_MAX_CONV, STAGE_E_UNSUPPORTED, "encoder: n_conv=%d", a->n_conv);
  STG_REQUIRE(a->out_mode == 0 || ((a->out_mode == 1 || a->out_mode == 2) && a->mask), STAGE_E_BADARG, "encoder: out_mode=%d needs a mask", a->out_mode);
  STG_REQUIRE(a->out_mode != 2 || (a->L > 64 && a->L <= CqaCfg::LS_MAX && !(a->flags & STAGE_ENC_NO_FINAL_LN)), STAGE_E_UNSUPPORTED,
              "encoder: out_mode 2 needs 64 < L <= %d (one sequence per tile) and the closing LayerNorm", CqaCfg::LS_MAX);
  STG_REQUIRE(aligned16(a->x) && aligned16(a->y) && (!a->pe || aligned16(a->pe)), STAGE_E_BADARG, "encoder: pointers must be 16-byte aligned");
  const int S = stage_blocked_stride(a->L);
  EncDev d;
  d.S = S;
  d.x = a->x; d.mask = a->mask; d.pe = a->pe; d.fin_g = a->final_ln_g; d.fin_b = a->final_ln_b; d.y = a->y;
  d.n_seq = a->n_seq; d.L = a->L; d.n_conv = a->n_conv; d.out_mode = a->out_mode; d.flags = a->flags; d.timeline = a->timeline;
  d.blob_bytes = CqaCfg::CTX_BYTES; d.blob_lbo = CqaCfg::LBO_CS; d.blob_part = CqaCfg::CS_PART;
  d.blob_inv_off = CqaCfg::CTX_CS_INV; d.blob_mask_off = CqaCfg::CTX_CS_MASK;
  for (int i = 0; i < STAGE_MAX_CONV; ++i) {
    const bool on = i < a->n_conv;
    if (on) STG_REQUIRE((a->ln_g[i] == nullptr) == (a->ln_b[i] == nullptr) && a->dw_w[i] && a->dw_b[i] && a->pw_packed[i] && a->pw_b[i], STAGE_E_BADARG,
                        "encoder: conv %d has a null weight", i);
    d.ln_g[i] = on ? a->ln_g[i] : nullptr; d.ln_b[i] = on ? a->ln_b[i] : nullptr;
    d.dw_w[i] = on ? a->dw_w[i] : nullptr; d.dw_b[i] = on ? a->dw_b[i] : nullptr;
    d.pw[i] = on ? static_cast<const uint8_t*>(a->pw_packed[i]) : nullptr; d.pw_b[i] = on ? a->pw_b[i] : nullptr;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (a->ksize) {
    case 3: return launch_encoder<3>(d, s);
    case 5: return launch_encoder<5>(d, s);
    case 7: return launch_encoder<7>(d, s);
    default: return fail(STAGE_E_UNSUPPORTED, "encoder: kernel_size=%d (built: 3, 5, 7)", a->ksize);
  }
}

int stage_mha_fwd(const StageMhaArgs* a, void* stream) {
  STG_REQUIRE(a && a->x && a->y, STAGE_E_BADARG, "mha: null pointer");
  STG_REQUIRE(a->n_seq > 0 && a->L > 0, STAGE_E_BADARG, "mha: n_seq=%lld L=%d", (long long)a->n_seq, a->L);
  STG_REQUIRE(a->L <= STAGE_MAX_SEQ, STAGE_E_UNSUPPORTED, "mha: L=%d > %d", a->L, STAGE_MAX_SEQ);
  STG_REQUIRE(a->n_heads == 1 || a->n_heads == 2 || a->n_heads == 4, STAGE_E_UNSUPPORTED, "mha: n_heads=%d (built: 1, 2, 4)", a->n_heads);
  STG_REQUIRE((a->ln_in_g == nullptr) == (a->ln_in_b == nullptr) && (a->final_ln_g == nullptr) == (a->final_ln_b == nullptr),
              STAGE_E_BADARG, "mha: LayerNorm gamma/beta must come in pairs");
  STG_REQUIRE(a->out_mode == 0 || (a->out_mode == 1 && a->mask), STAGE_E_BADARG, "mha: out_mode=%d needs a mask", a->out_mode);
  STG_REQUIRE(aligned16(a->x) && aligned16(a->y), STAGE_E_BADARG, "mha: pointers must be 16-byte aligned");
  MhaDev d;
  d.x = a->x; d.mask = a->mask; d.ln_g = a->ln_in_g; d.ln_b = a->ln_in_b; d.fin_g = a->final_ln_g; d.fin_b = a->final_ln_b;
  d.y = a->y; d.n_seq = a->n_seq; d.L = a->L; d.n_heads = a->n_heads; d.residual = a->residual; d.out_mode = a->out_mode;
  for (int i = 0; i < 4; ++i) {
    STG_REQUIRE(a->w_packed[i] && a->b[i] && aligned16(a->w_packed[i]) && aligned16(a->b[i]), STAGE_E_BADARG,
                "mha: linears.%d weight/bias null or misaligned", i);
    d.w[i] = static_cast<const uint8_t*>(a->w_packed[i]);
    d.b[i] = a->b[i];
  }
  int rc = set_smem(mha_kernel, MhaCfg::SMEM_BYTES, "mha");
  if (rc) return rc;
  const int G = 128 / a->L;
  const long long tiles = (a->n_seq + G - 1) / G;
  const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
  mha_kernel<<<grid, MhaCfg::THREADS, MhaCfg::SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(d);
  return check_launch("mha");
}

// weight blob layout of the fused CQA kernel
static constexpr size_t WB_WA = 0;                      // (Wc*gamma)[:, 0:128]   packed kc=64 (Za = a . Wa^T via the linear kernel)
static constexpr size_t WB_WU = 65536;                  // (Wc*gamma): [aU h0 | U h0 | aU h1 | U h1] in the kernel's K order, kc=32, 2 chunks each
static constexpr size_t WB_WF = WB_WU + 8 * 16384;      // (Wf*gamma)             packed kc=32, 12 chunks
static constexpr size_t WB_VEC = WB_WF + 12 * 16384;    // c2q cs | c2q cb | cat cs | cat cb | zeros, 128 floats each
static constexpr size_t WB_BYTES = WB_VEC + 5 * 512;

size_t stage_cqa_weight_blob_bytes(void) { return WB_BYTES; }

int stage_cqa_prepare_weights(const float* c2q_ln_g, const float* c2q_ln_b, const float* c2q_w, const float* c2q_b,
                              const float* cat_ln_g, const float* cat_ln_b, const float* cat_w, const float* cat_b,
                              void* wblob, void* stream) {
  STG_REQUIRE(c2q_ln_g && c2q_ln_b && c2q_w && c2q_b && cat_ln_g && cat_ln_b && cat_w && cat_b && wblob, STAGE_E_BADARG,
              "cqa_prepare_weights: null pointer");
  STG_REQUIRE(aligned16(wblob), STAGE_E_BADARG, "cqa_prepare_weights: blob must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* wb = static_cast<uint8_t*>(wblob);
  float* vec = reinterpret_cast<float*>(wb + WB_VEC);
  {
    PackGeom g = pack_geom(128, 128, 64);
    pack_weight_kernel<<<((g.kpad / 8) * g.npad + 255) / 256, 256, 0, s>>>(c2q_w, 384, 0, c2q_ln_g, g, wb + WB_WA);
  }
  {  // c2q chunks in the order the fused kernel consumes them.  Both segments use the kernel's K order (perm128: packed
     // position h*64 + q*16 + j <- column q*32 + h*16 + j): per half h, the two a*U chunks (columns 256..383) and then
     // the two U chunks (columns 128..255)  ->  blob = [aU h0 | U h0 | aU h1 | U h1], two chunks each
    PackGeom gh = pack_geom(128, 64, 32);  // one half (h) of a permuted segment = 2 chunks = packed positions [64h, 64h+64)
    for (int h = 0; h < 2; ++h) {
      pack_weight_half_kernel<<<((gh.kpad / 8) * gh.npad + 255) / 256, 256, 0, s>>>(c2q_w, 384, 256, c2q_ln_g, gh, h, wb + WB_WU + (4 * h) * 16384);
      pack_weight_half_kernel<<<((gh.kpad / 8) * gh.npad + 255) / 256, 256, 0, s>>>(c2q_w, 384, 128, c2q_ln_g, gh, h, wb + WB_WU + (4 * h + 2) * 16384);
    }
  }
  {
    PackGeom g = pack_geom(128, 384, 32);
    pack_weight_kernel<<<((g.kpad / 8) * g.npad + 255) / 256, 256, 0, s>>>(cat_w, 384, 0, cat_ln_g, g, wb + WB_WF, 0);
  }
  fold_layernorm_kernel<<<16, 256, 0, s>>>(c2q_w, 128, 384, c2q_ln_g, c2q_ln_b, c2q_b, vec, vec + 128);
  fold_layernorm_kernel<<<16, 256, 0, s>>>(cat_w, 128, 384, cat_ln_g, cat_ln_b, cat_b, vec + 256, vec + 384);
  cudaMemsetAsync(vec + 512, 0, 512, s);
  return check_launch("cqa_prepare_weights");
}

static size_t align128(size_t x) { return (x + 127) & ~static_cast<size_t>(127); }

size_t stage_cqa_workspace_bytes(int n, int t, int lq) {
  if (n <= 0 || t <= 0 || lq <= 0) return 0;
  const size_t n_pt = (5 * static_cast<size_t>(lq) + 127) / 128;
  return align128(static_cast<size_t>(n) * n_pt * CqaCfg::ABLOB_BYTES) + align128(static_cast<size_t>(n) * t * CqaCfg::CTX_BYTES) +
         2 * align128(static_cast<size_t>(n) * 5 * lq * 128 * sizeof(float));  // a.Wa and the blocked fp32 answers
}

void* stage_cqa_ctx_blobs(void* workspace, int n, int t, int lq) {
  if (!workspace || n <= 0 || t <= 0 || lq <= 0) return nullptr;
  const size_t n_pt = (5 * static_cast<size_t>(lq) + 127) / 128;
  return static_cast<uint8_t*>(workspace) + align128(static_cast<size_t>(n) * n_pt * CqaCfg::ABLOB_BYTES);
}
size_t stage_cqa_ctx_blob_bytes(void) { return CqaCfg::CTX_BYTES; }

int stage_cqa_fused_fwd(const StageCqaArgs* a, void* stream) {
  STG_REQUIRE(a && a->a_embed && a->a_mask && (a->sub_embed || a->sub_prepared) && a->sub_mask && a->vid_embed && a->vid_mask, STAGE_E_BADARG,
              "cqa: null input (single-stream models are not built yet)");
  STG_REQUIRE(a->wblob && a->cat_out_ln_g && a->cat_out_ln_b && a->workspace, STAGE_E_BADARG, "cqa: null weight blob / workspace");
  STG_REQUIRE(a->statement && a->statement_mask, STAGE_E_BADARG, "cqa: null output");
  STG_REQUIRE(a->n > 0 && a->t > 0 && a->lq > 0 && a->ls > 0 && a->r > 0, STAGE_E_BADARG, "cqa: non-positive size");
  STG_REQUIRE(a->lq <= 128 && a->ls <= CqaCfg::LS_MAX && a->r <= CqaCfg::R_MAX, STAGE_E_UNSUPPORTED,
              "cqa: lq=%d ls=%d r=%d exceed the tile (128, %d, %d)", a->lq, a->ls, a->r, CqaCfg::LS_MAX, CqaCfg::R_MAX);
  STG_REQUIRE(aligned16(a->a_embed) && (!a->sub_embed || aligned16(a->sub_embed)) && aligned16(a->vid_embed) && aligned16(a->statement) &&
                  aligned16(a->wblob) && (reinterpret_cast<uintptr_t>(a->workspace) & 127u) == 0, STAGE_E_BADARG,
              "cqa: pointers must be 16-byte (workspace 128-byte) aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const uint8_t* wb = static_cast<const uint8_t*>(a->wblob);
  const float* vec = reinterpret_cast<const float*>(wb + WB_VEC);
  uint8_t* ws = static_cast<uint8_t*>(a->workspace);
  uint8_t* a_blob = ws;
  const int n_pt = (5 * a->lq + 127) / 128;
  uint8_t* ctx_blob = a_blob + align128(static_cast<size_t>(a->n) * n_pt * CqaCfg::ABLOB_BYTES);
  float* za = reinterpret_cast<float*>(ctx_blob + align128(static_cast<size_t>(a->n) * a->t * CqaCfg::CTX_BYTES));
  float* a_blk = za + align128(static_cast<size_t>(a->n) * 5 * a->lq * 128 * sizeof(float)) / sizeof(float);
  const int ls_pad = round_up(a->ls, 16), r_pad = round_up(a->r, 16);
  // 1. operand images
  {
    CqaPrepDev d;
    d.n = a->n; d.t = a->t; d.lq = a->lq; d.ls = a->ls; d.r = a->r; d.ls_pad = ls_pad; d.r_pad = r_pad; d.n_pt = n_pt;
    d.a_embed = a->a_embed; d.a_mask = a->a_mask; d.sub_embed = a->sub_embed; d.sub_mask = a->sub_mask;
    d.vid_embed = a->vid_embed; d.vid_mask = a->vid_mask; d.a_blob = a_blob; d.ctx_blob = ctx_blob; d.a_blk = a_blk;
    d.sub_prepared = a->sub_prepared;
    const long long warps = static_cast<long long>(a->n) * n_pt * 128 + static_cast<long long>(a->n) * a->t * ((a->sub_prepared ? 0 : ls_pad) + r_pad);
    cqa_prep_kernel<<<static_cast<unsigned>((warps + 7) / 8), 256, 0, s>>>(d);
    int rc = check_launch("cqa_prep");
    if (rc) return rc;
  }
  // 2. frame-invariant third of c2q: Za = a . (Wc*gamma)[:, :128]^T
  {
    LinearDev d;
    d.x = a->a_embed; d.w = wb + WB_WA; d.bias = vec + 512; d.cs = nullptr; d.cb = nullptr;
    d.ln_out_g = nullptr; d.ln_out_b = nullptr; d.y = za; d.m = static_cast<long long>(a->n) * 5 * a->lq;
    d.k = 128; d.n = 128; d.n_chunks = 2; d.l2norm_in = 0; d.relu = 0; d.out_blocked = 1; d.rows_per_block = 5 * a->lq; d.timeline = nullptr;
    d.ldy = 128; d.x_rowmax = nullptr;
    int rc = launch_linear<128, 64, 2>(d, s);
    if (rc) return rc;
  }
  // 3. the fused kernel
  CqaDev d;
  d.n = a->n; d.t = a->t; d.lq = a->lq; d.ls = a->ls; d.r = a->r; d.ls_pad = ls_pad; d.r_pad = r_pad; d.n_pt = n_pt; d.scale = a->scale;
  d.a_blk = a_blk; d.a_blob = a_blob; d.ctx_blob = ctx_blob; d.za = za;
  d.c2q_w = wb + WB_WU; d.c2q_cs = vec; d.c2q_cb = vec + 128; d.cat_w = wb + WB_WF; d.cat_cs = vec + 256; d.cat_cb = vec + 384;
  d.out_g = a->cat_out_ln_g; d.out_b = a->cat_out_ln_b;
  d.statement = a->statement; d.statement_mask = a->statement_mask;
  d.stmt_stride = a->statement_blocked ? stage_blocked_stride(a->lq) : 0;
  d.vid_raw_s = a->vid_raw_s; d.sub_raw_s = a->sub_raw_s; d.vid_norm_s = a->vid_norm_s; d.sub_norm_s = a->sub_norm_s;
  d.att_vid = a->att_vid; d.att_sub = a->att_sub; d.mixed_vid = a->mixed_vid; d.mixed_sub = a->mixed_sub;
  d.timeline = a->timeline;
  const int frames = a->n * a->t;
  const int grid = frames < num_sms() ? frames : num_sms();
  if (a->timeline != nullptr) {  // instrumented build of the same kernel (tools/cqa_timeline.py)
    int rc = set_smem(cqa_kernel<true>, CqaCfg::SMEM_BYTES, "cqa");
    if (rc) return rc;
    cqa_kernel<true><<<grid, CqaCfg::THREADS, CqaCfg::SMEM_BYTES, s>>>(d);
  } else {
    int rc = set_smem(cqa_kernel<false>, CqaCfg::SMEM_BYTES, "cqa");
    if (rc) return rc;
    cqa_kernel<false><<<grid, CqaCfg::THREADS, CqaCfg::SMEM_BYTES, s>>>(d);
  }
  return check_launch("cqa");
}

int stage_span_head_fwd(const StageHeadArgs* a, void* stream) {
  STG_REQUIRE(a && a->max_statement && a->statement_mask && a->ts_mask && a->t_scores && a->t_probs && a->spans &&
                  a->span_conf && a->answer && a->workspace, STAGE_E_BADARG, "span_head: null pointer");
  STG_REQUIRE(a->proj_ln_g && a->proj_ln_b && a->proj_w && a->proj_b && a->st_ln_g && a->st_ln_b && a->st_w && a->st_b &&
                  a->ed_ln_g && a->ed_ln_b && a->ed_w && a->ed_b && a->cls_ln_g && a->cls_ln_b && a->cls_w && a->cls_b,
              STAGE_E_BADARG, "span_head: null weight");
  STG_REQUIRE(a->n > 0 && a->t > 0 && a->lq > 0, STAGE_E_BADARG, "span_head: n=%d t=%d lq=%d", a->n, a->t, a->lq);
  STG_REQUIRE(a->t <= 8192, STAGE_E_UNSUPPORTED, "span_head: t=%d > 8192", a->t);
  STG_REQUIRE(aligned16(a->max_statement) && aligned16(a->workspace), STAGE_E_BADARG, "span_head: pointers must be 16-byte aligned");
  HeadDev d;
  d.n = a->n; d.t = a->t; d.lq = a->lq; d.add_local = a->add_local; d.extra = a->extra_span_length;
  d.max_statement = a->max_statement; d.statement_mask = a->statement_mask; d.ts_mask = a->ts_mask;
  d.proj_g = a->proj_ln_g; d.proj_b = a->proj_ln_b; d.proj_w = a->proj_w; d.proj_bias = a->proj_b;
  d.st_g = a->st_ln_g; d.st_b = a->st_ln_b; d.st_w = a->st_w; d.st_bias = a->st_b;
  d.ed_g = a->ed_ln_g; d.ed_b = a->ed_ln_b; d.ed_w = a->ed_w; d.ed_bias = a->ed_b;
  d.cls_g = a->cls_ln_g; d.cls_b = a->cls_ln_b; d.cls_w = a->cls_w; d.cls_bias = a->cls_b;
  d.t_scores = a->t_scores; d.t_probs = a->t_probs; d.spans = a->spans; d.span_conf = a->span_conf; d.answer = a->answer;
  d.enc = a->workspace;
  STG_REQUIRE(a->ts_terms == nullptr || (a->target && a->ts_st && a->ts_ed && a->ts_loss), STAGE_E_BADARG,
              "span_head: the temporal loss needs target, ts_st, ts_ed and ts_loss");
  d.target = reinterpret_cast<const long long*>(a->target); d.ts_st = reinterpret_cast<const long long*>(a->ts_st);
  d.ts_ed = reinterpret_cast<const long long*>(a->ts_ed); d.ts_terms = a->ts_terms; d.frame_mask = a->frame_mask; d.gfeat = a->gfeat;
  const size_t smem = (128 * 128 + 16 * 128 + 3 * static_cast<size_t>(a->t) + 256) * sizeof(float);
  int rc = set_smem(head_kernel, smem, "span_head");
  if (rc) return rc;
  head_kernel<<<a->n * 5, 256, smem, static_cast<cudaStream_t>(stream)>>>(d);
  if (a->ts_terms != nullptr) ts_loss_sum_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(a->ts_terms, a->n, a->ts_loss);
  return check_launch("span_head");
}

int stage_train_proposals_fwd(const StageProposalArgs* a, void* stream) {
  STG_REQUIRE(a && a->encoded && a->frame_mask && a->gfeat && a->spans && a->span_conf && a->target && a->ts_st && a->ts_ed &&
                  a->cls_ln_g && a->cls_ln_b && a->cls_w && a->cls_b && a->answer_rows && a->row_valid && a->row_target,
              STAGE_E_BADARG, "train_proposals: null pointer");
  STG_REQUIRE(a->n > 0 && a->t > 0 && a->extra_span_length >= 0, STAGE_E_BADARG, "train_proposals: n=%d t=%d", a->n, a->t);
  ProposalDev d;
  d.n = a->n; d.t = a->t; d.extra = a->extra_span_length; d.ce_prob_thd = a->ce_prob_thd;
  d.enc = a->encoded; d.frame_mask = a->frame_mask; d.gfeat = a->gfeat; d.spans = a->spans; d.span_conf = a->span_conf;
  d.target = reinterpret_cast<const long long*>(a->target); d.ts_st = reinterpret_cast<const long long*>(a->ts_st);
  d.ts_ed = reinterpret_cast<const long long*>(a->ts_ed);
  d.cls_g = a->cls_ln_g; d.cls_b = a->cls_ln_b; d.cls_w = a->cls_w; d.cls_bias = a->cls_b;
  d.answer_rows = a->answer_rows; d.row_valid = a->row_valid; d.row_target = reinterpret_cast<long long*>(a->row_target);
  proposal_cls_kernel<<<2 * a->n, 256, 0, static_cast<cudaStream_t>(stream)>>>(d);
  return check_launch("train_proposals");
}

int stage_find_max_span(const float* p_st, const float* p_ed, int rows, int t, int32_t* spans, float* conf, void* stream) {
  STG_REQUIRE(p_st && p_ed && spans && conf, STAGE_E_BADARG, "find_max_span: null pointer");
  STG_REQUIRE(rows > 0 && t > 0, STAGE_E_BADARG, "find_max_span: rows=%d t=%d", rows, t);
  find_max_span_kernel<<<(rows + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(p_st, p_ed, rows, t, spans, conf);
  return check_launch("find_max_span");
}

int stage_find_max_pair(const float* p1, const float* p2, int rows, int t, int64_t row_stride, int elem_stride,
                        int32_t* spans, double* val, void* stream) {
  STG_REQUIRE(p1 && p2 && spans && val, STAGE_E_BADARG, "find_max_pair: null pointer");
  STG_REQUIRE(rows > 0 && t > 0 && row_stride > 0 && elem_stride > 0, STAGE_E_BADARG, "find_max_pair: rows=%d t=%d", rows, t);
  find_max_pair_kernel<<<(rows + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(p1, p2, rows, t, row_stride, elem_stride, spans, val);
  return check_launch("find_max_pair");
}

int stage_att_loss_fwd(const float* scores, const int64_t* pos_off, const int64_t* neg_off, int64_t m, int loss_type,
                       float alpha, float margin, float* partials, float* loss, void* stream) {
  STG_REQUIRE(scores && pos_off && neg_off && partials && loss, STAGE_E_BADARG, "att_loss: null pointer");
  STG_REQUIRE(m > 0 && m < (1ll << 31) * 256, STAGE_E_BADARG, "att_loss: m=%lld", (long long)m);
  STG_REQUIRE(loss_type == 0 || loss_type == 1, STAGE_E_UNSUPPORTED, "att_loss: loss_type=%d (0 = lse, 1 = hinge)", loss_type);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int blocks = static_cast<int>((m + 255) / 256);
  att_loss_kernel<<<blocks, 256, 0, s>>>(scores, reinterpret_cast<const long long*>(pos_off), reinterpret_cast<const long long*>(neg_off),
                                        m, loss_type, alpha, margin, partials);
  sum_in_order_kernel<<<1, 32, 0, s>>>(partials, blocks, loss);
  return check_launch("att_loss");
}

int stage_pad_rows(const float* rows, const int64_t* off, const int32_t* len, int64_t n_seq, int l_pad, int d, float* out,
                   float* mask, void* stream) {
  STG_REQUIRE(off && len && out && mask, STAGE_E_BADARG, "pad_rows: null pointer");
  STG_REQUIRE(n_seq > 0 && l_pad > 0 && d > 0 && d % 4 == 0, STAGE_E_BADARG, "pad_rows: n_seq=%lld l_pad=%d d=%d", (long long)n_seq, l_pad, d);
  STG_REQUIRE((rows == nullptr || aligned16(rows)) && aligned16(out), STAGE_E_BADARG, "pad_rows: pointers must be 16-byte aligned");
  const long long n_rows = n_seq * l_pad;
  const long long want = (n_rows + 7) / 8;  // 8 warps per block
  const int grid = static_cast<int>(want < 148ll * 32 ? want : 148ll * 32);
  pad_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(rows, reinterpret_cast<const long long*>(off), len, n_seq, l_pad, d / 4, out, mask);
  return check_launch("pad_rows");
}

int stage_mask_logits(const float* x, const float* mask, int64_t outer, int inner, float* y, void* stream) {
  STG_REQUIRE(x && mask && y, STAGE_E_BADARG, "mask_logits: null pointer");
  STG_REQUIRE(outer > 0 && inner > 0, STAGE_E_BADARG, "mask_logits: outer=%lld inner=%d", (long long)outer, inner);
  const long long total = outer * inner;
  const int grid = static_cast<int>((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  mask_logits_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, mask, outer, inner, y);
  return check_launch("mask_logits");
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
