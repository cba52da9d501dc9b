// extern "C" entry points of libnnab.so — see include/nnab.h for the contract
// and the reference file:line each call replaces.
#include <atomic>
#include <mutex>
#include <string.h>
#include <utility>
#include <vector>

#include "common.cuh"

namespace nnab {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_cuda_error(const char* where, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}
void set_error_text(const char* text) { snprintf(g_err, sizeof(g_err), "%s", text); }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// T of a centred / un-centred framing: (L + 2*pad - K)/hop + 1
static inline int64_t frames_of(int64_t L, int K, int hop, int pad) {
  const int64_t span = L + 2 * (int64_t)pad - K;
  return span < 0 ? 0 : span / hop + 1;
}

static int check_common(const float* x, int64_t B, int64_t L, int64_t x_pitch, int K, int F,
                        int hop, int pad, int pad_mode, int64_t T) {
  if (x == nullptr || B < 0 || L <= 0 || x_pitch < L || K <= 0 || F <= 0 || hop <= 0)
    return NNAB_EINVAL;
  if (pad_mode != NNAB_PAD_REFLECT && pad_mode != NNAB_PAD_CONSTANT) return NNAB_EINVAL;
  // nn.ReflectionPad1d needs pad < L; callers raise the reference's exception first.
  if (pad > 0 && pad_mode == NNAB_PAD_REFLECT && pad >= L) return NNAB_EINVAL;
  if (T != frames_of(L, K, hop, pad) || T <= 0) return NNAB_EINVAL;
  return NNAB_OK;
}

static int check_arch() {
  int dev = 0;
  NNAB_CUDA_TRY(cudaGetDevice(&dev));
  int major = 0;
  NNAB_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  return major == 10 ? NNAB_OK : NNAB_EARCH;
}

// ---- optional in-stream timing of the framed contraction (bench.py) --------
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_pairs;  // recorded, unread
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_free;

static bool prof_begin(cudaStream_t s, std::pair<cudaEvent_t, cudaEvent_t>* pr) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return false;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof_free.empty()) {
    *pr = g_prof_free.back();
    g_prof_free.pop_back();
  } else {
    if (cudaEventCreate(&pr->first) != cudaSuccess) return false;
    if (cudaEventCreate(&pr->second) != cudaSuccess) return false;
  }
  cudaEventRecord(pr->first, s);
  return true;
}
static void prof_end(cudaStream_t s, const std::pair<cudaEvent_t, cudaEvent_t>& pr) {
  cudaEventRecord(pr.second, s);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_pairs.push_back(pr);
}

static int run_framed_inner(const FramedProblem& p, const void* packed, void* ws, size_t ws_bytes,
                            int path, cudaStream_t stream);

// Run one framed contraction on the requested kernel family.
static int run_framed(const FramedProblem& p, const void* packed, void* ws, size_t ws_bytes,
                      int path, cudaStream_t stream) {
  std::pair<cudaEvent_t, cudaEvent_t> pr;
  const bool timed = prof_begin(stream, &pr);
  const int rc = run_framed_inner(p, packed, ws, ws_bytes, path, stream);
  if (timed) prof_end(stream, pr);
  return rc;
}

static int run_framed_inner(const FramedProblem& p, const void* packed, void* ws, size_t ws_bytes,
                      int path, cudaStream_t stream) {
  bool use_tc = false;
  if (path == NNAB_PATH_TCGEN05) {
    if (packed == nullptr || !tc_supported(p)) return NNAB_EALIGN;
    use_tc = true;
  } else if (path == NNAB_PATH_AUTO) {
    use_tc = (packed != nullptr) && tc_supported(p);
  } else if (path != NNAB_PATH_SIMT) {
    return NNAB_EINVAL;
  }
  if (use_tc) return launch_framed_tc(p, packed, ws, ws_bytes, stream);
  return launch_framed_simt(p, stream);
}

static bool wants_tc(int path, int K, int hop) {
  if (path == NNAB_PATH_SIMT) return false;
  FramedProblem q{};
  q.K = K; q.hop = hop; q.F = 1; q.B = 1; q.L = K; q.T = 1;
  return tc_supported(q);
}

}  // namespace nnab

using namespace nnab;

extern "C" {

int nnab_abi_version(void) { return NNAB_ABI_VERSION; }

const char* nnab_strerror(int status) {
  switch (status) {
    case NNAB_OK: return "ok";
    case NNAB_EINVAL: return "invalid argument or shape mismatch";
    case NNAB_EALIGN: return "tcgen05 path forced but shape/alignment rules not met";
    case NNAB_EARCH: return "device is not sm_100 (B200)";
    case NNAB_ECUDA: return "CUDA error";
    case NNAB_EWORKSPACE: return "workspace missing or too small";
    case NNAB_EUNSUPPORTED: return "unsupported size";
    default: return "unknown status";
  }
}

const char* nnab_last_cuda_error(void) { return g_err; }

uint64_t nnab_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void nnab_profile_enable(int on) { g_prof_on.store(on ? 1 : 0); }

int nnab_profile_read(double* framed_ms, uint64_t* framed_launches) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double total = 0.0;
  uint64_t n = 0;
  for (auto& pr : g_prof_pairs) {
    NNAB_CUDA_TRY(cudaEventSynchronize(pr.second));
    float ms = 0.f;
    NNAB_CUDA_TRY(cudaEventElapsedTime(&ms, pr.first, pr.second));
    total += ms;
    ++n;
    g_prof_free.push_back(pr);
  }
  g_prof_pairs.clear();
  if (framed_ms) *framed_ms = total;
  if (framed_launches) *framed_launches = n;
  return NNAB_OK;
}

int nnab_pack_tile_n(void) { return tc_tile_n(); }
size_t nnab_packed_basis_bytes(int F, int K) { return tc_packed_bytes(F, K); }
int nnab_pack_basis(const float* w_re, const float* w_im, int F, int K, void* packed,
                    void* stream) {
  if (w_re == nullptr || w_im == nullptr || packed == nullptr || F <= 0 || K <= 0)
    return NNAB_EINVAL;
  return tc_pack_basis(w_re, w_im, F, K, packed, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ STFT ----
size_t nnab_stft_workspace_bytes(int64_t B, int64_t L, int n_fft, int F, int hop, int center,
                                 int path) {
  (void)F;
  if (!wants_tc(path, n_fft, hop)) return 0;
  return tc_workspace_bytes(B, L, n_fft, hop, center ? n_fft / 2 : 0);
}

int nnab_stft_forward(const float* x, int64_t B, int64_t L, int64_t x_pitch, const float* wcos,
                      const float* wsin, const void* packed, int n_fft, int F, int hop,
                      int center, int pad_mode, int out_format, float sqrt_eps, float* out,
                      int64_t T, void* workspace, size_t ws_bytes, int path, void* stream) {
  const int pad = center ? n_fft / 2 : 0;
  int rc = check_common(x, B, L, x_pitch, n_fft, F, hop, pad, pad_mode, T);
  if (rc) return rc;
  if (wcos == nullptr || wsin == nullptr || out == nullptr) return NNAB_EINVAL;
  if (out_format != NNAB_FMT_MAGNITUDE && out_format != NNAB_FMT_COMPLEX &&
      out_format != NNAB_FMT_PHASE_ANGLE)
    return NNAB_EINVAL;
  if ((rc = check_arch())) return rc;
  FramedProblem p{};
  p.x = x; p.B = B; p.L = L; p.x_pitch = x_pitch;
  p.w_re = wcos; p.w_im = wsin; p.F = F; p.K = n_fft; p.hop = hop;
  p.pad = pad; p.pad_mode = pad_mode; p.scale = nullptr; p.scale_all = 1.f;
  p.fmt = out_format; p.eps = sqrt_eps; p.power = 1.f; p.out = out; p.T = T;
  p.out_bins = F; p.bin_offset = 0;
  return run_framed(p, packed, workspace, ws_bytes, path, (cudaStream_t)stream);
}

// ------------------------------------------------- Mel / Gammatone / MFCC ----
// workspace layout: [P (B,F,T) fp32][tc scratch]  (+ [mel (B,n_mels,T)][B words] for MFCC)
static size_t power_bytes(int64_t B, int F, int64_t T) {
  return align_up((size_t)B * F * T * sizeof(float), 256);
}

size_t nnab_filterbank_table_bytes(int F) { return (size_t)F * sizeof(FbEntry) + 64; }

int nnab_build_filterbank_table(const float* fb, int n_fb, int F, void* table, int* h_max_nnz,
                                void* stream) {
  if (fb == nullptr || table == nullptr || h_max_nnz == nullptr || n_fb <= 0 || F <= 0)
    return NNAB_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  int* d_max = reinterpret_cast<int*>(reinterpret_cast<char*>(table) + (size_t)F * sizeof(FbEntry));
  int rc = launch_fb_table(fb, n_fb, F, reinterpret_cast<FbEntry*>(table), d_max, s);
  if (rc) return rc;
  NNAB_CUDA_TRY(cudaMemcpyAsync(h_max_nnz, d_max, sizeof(int), cudaMemcpyDeviceToHost, s));
  NNAB_CUDA_TRY(cudaStreamSynchronize(s));
  return NNAB_OK;
}

static bool fused_fbank(int path, const void* packed, const void* fb_table, int n_fft, int hop) {
  return fb_table != nullptr && packed != nullptr && path != NNAB_PATH_SIMT &&
         wants_tc(path, n_fft, hop);
}

size_t nnab_filterbank_workspace_bytes(int64_t B, int64_t L, int n_fft, int F, int hop,
                                       int center, int n_fb, int path, int has_table) {
  (void)n_fb;
  const int pad = center ? n_fft / 2 : 0;
  const int64_t T = frames_of(L, n_fft, hop, pad);
  const bool tc = wants_tc(path, n_fft, hop);
  size_t n = 0;
  if (!(has_table && tc)) n += power_bytes(B, F, T);  // un-fused: (B,F,T) power spectrogram
  if (tc) n += tc_workspace_bytes(B, L, n_fft, hop, pad);
  return n;
}

static int power_spectrogram(const float* x, int64_t B, int64_t L, int64_t x_pitch,
                             const float* wcos, const float* wsin, const void* packed,
                             int n_fft, int F, int hop, int pad, int pad_mode, float sqrt_eps,
                             float power, float* P, int64_t T, void* tc_ws, size_t tc_ws_bytes,
                             int path, cudaStream_t stream) {
  FramedProblem p{};
  p.x = x; p.B = B; p.L = L; p.x_pitch = x_pitch;
  p.w_re = wcos; p.w_im = wsin; p.F = F; p.K = n_fft; p.hop = hop;
  p.pad = pad; p.pad_mode = pad_mode; p.scale = nullptr; p.scale_all = 1.f;
  p.fmt = FMT_POWER; p.eps = sqrt_eps; p.power = power; p.out = P; p.T = T;
  p.out_bins = F; p.bin_offset = 0;
  return run_framed(p, packed, tc_ws, tc_ws_bytes, path, stream);
}

int nnab_stft_filterbank_forward(const float* x, int64_t B, int64_t L, int64_t x_pitch,
                                 const float* wcos, const float* wsin, const void* packed,
                                 int n_fft, int F, int hop, int center, int pad_mode,
                                 float sqrt_eps, float power, const float* fb, int n_fb,
                                 const void* fb_table, float* out, int64_t T, void* workspace,
                                 size_t ws_bytes, int path, void* stream) {
  const int pad = center ? n_fft / 2 : 0;
  int rc = check_common(x, B, L, x_pitch, n_fft, F, hop, pad, pad_mode, T);
  if (rc) return rc;
  if (wcos == nullptr || wsin == nullptr || fb == nullptr || out == nullptr || n_fb <= 0)
    return NNAB_EINVAL;
  if ((rc = check_arch())) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  bool fused = fused_fbank(path, packed, fb_table, n_fft, hop);
  if (fused) {
    FramedProblem p{};
    p.x = x; p.B = B; p.L = L; p.x_pitch = x_pitch;
    p.w_re = wcos; p.w_im = wsin; p.F = F; p.K = n_fft; p.hop = hop;
    p.pad = pad; p.pad_mode = pad_mode; p.scale = nullptr; p.scale_all = 1.f;
    p.fmt = FMT_FBANK; p.eps = sqrt_eps; p.power = power; p.out = out; p.T = T;
    p.out_bins = n_fb; p.bin_offset = 0;
    p.fb_table = reinterpret_cast<const FbEntry*>(fb_table); p.n_fb = n_fb;
    if (tc_supported(p)) {
      const size_t need = tc_workspace_bytes(B, L, n_fft, hop, pad);
      if (workspace == nullptr || ws_bytes < need) return NNAB_EWORKSPACE;
      // the epilogue accumulates filter sums with fp32 atomics: start from zero
      NNAB_CUDA_TRY(cudaMemsetAsync(out, 0, (size_t)B * n_fb * T * sizeof(float), s));
      return run_framed(p, packed, workspace, ws_bytes, NNAB_PATH_TCGEN05, s);
    }
    fused = false;
  }
  const size_t need = nnab_filterbank_workspace_bytes(B, L, n_fft, F, hop, center, n_fb, path, 0);
  if (workspace == nullptr || ws_bytes < need) return NNAB_EWORKSPACE;
  float* P = (float*)workspace;
  const size_t pb = power_bytes(B, F, T);
  rc = power_spectrogram(x, B, L, x_pitch, wcos, wsin, packed, n_fft, F, hop, pad, pad_mode,
                         sqrt_eps, power, P, T, (char*)workspace + pb, ws_bytes - pb, path, s);
  if (rc) return rc;
  return launch_filterbank(P, fb, B, F, T, n_fb, out, s);
}

static size_t mel_bytes(int64_t B, int n_mels, int64_t T) {
  return align_up((size_t)B * n_mels * T * sizeof(float), 256);
}

size_t nnab_mfcc_workspace_bytes(int64_t B, int64_t L, int n_fft, int F, int hop, int center,
                                 int n_mels, int path, int has_table) {
  const int pad = center ? n_fft / 2 : 0;
  const int64_t T = frames_of(L, n_fft, hop, pad);
  // the un-fused size is the upper bound (a huge batch can still fall back to it)
  const size_t fbw = nnab_filterbank_workspace_bytes(B, L, n_fft, F, hop, center, n_mels, path, 0);
  (void)has_table;
  return align_up(fbw, 256) + mel_bytes(B, n_mels, T) + align_up((size_t)B * sizeof(unsigned int), 256);
}

int nnab_mfcc_forward(const float* x, int64_t B, int64_t L, int64_t x_pitch, const float* wcos,
                      const float* wsin, const void* packed, int n_fft, int F, int hop,
                      int center, int pad_mode, float sqrt_eps, float power,
                      const float* mel_basis, int n_mels, const void* fb_table, float amin,
                      float ref, float top_db, const float* dct, int n_mfcc, float* out,
                      int64_t T, void* workspace, size_t ws_bytes, int path, void* stream) {
  const int pad = center ? n_fft / 2 : 0;
  int rc = check_common(x, B, L, x_pitch, n_fft, F, hop, pad, pad_mode, T);
  if (rc) return rc;
  if (wcos == nullptr || wsin == nullptr || mel_basis == nullptr || dct == nullptr ||
      out == nullptr || n_mels <= 0 || n_mfcc <= 0 || !(amin > 0.f))
    return NNAB_EINVAL;
  if ((rc = check_arch())) return rc;
  const size_t need = nnab_mfcc_workspace_bytes(B, L, n_fft, F, hop, center, n_mels, path, 0);
  if (workspace == nullptr || ws_bytes < need) return NNAB_EWORKSPACE;
  const size_t fbw =
      align_up(nnab_filterbank_workspace_bytes(B, L, n_fft, F, hop, center, n_mels, path, 0), 256);
  float* mel = (float*)((char*)workspace + fbw);
  unsigned int* scratch = (unsigned int*)((char*)workspace + fbw + mel_bytes(B, n_mels, T));
  rc = nnab_stft_filterbank_forward(x, B, L, x_pitch, wcos, wsin, packed, n_fft, F, hop, center,
                                    pad_mode, sqrt_eps, power, mel_basis, n_mels, fb_table, mel,
                                    T, workspace, fbw, path, stream);
  if (rc) return rc;
  return launch_mfcc_tail(mel, B, n_mels, T, amin, ref, top_db, dct, n_mfcc, out, scratch,
                          (cudaStream_t)stream);
}

// ------------------------------------------------------------- CQT1992v2 ----
size_t nnab_cqt1992v2_workspace_bytes(int64_t B, int64_t L, int width, int n_bins, int hop,
                                      int center, int path) {
  (void)n_bins;
  if (!wants_tc(path, width, hop)) return 0;
  return tc_workspace_bytes(B, L, width, hop, center ? width / 2 : 0);
}

int nnab_cqt1992v2_forward(const float* x, int64_t B, int64_t L, int64_t x_pitch,
                           const float* k_real, const float* k_imag, const void* packed,
                           const int32_t* h_k_begin, const int32_t* h_k_end, int n_bins,
                           int width, int hop, int center, int pad_mode, const float* scale,
                           float scale_all, int out_format, float sqrt_eps, float* out,
                           int64_t T, void* workspace, size_t ws_bytes, int path,
                           void* stream) {
  const int pad = center ? width / 2 : 0;
  int rc = check_common(x, B, L, x_pitch, width, n_bins, hop, pad, pad_mode, T);
  if (rc) return rc;
  if (k_real == nullptr || k_imag == nullptr || out == nullptr) return NNAB_EINVAL;
  if (out_format != NNAB_FMT_MAGNITUDE && out_format != NNAB_FMT_COMPLEX &&
      out_format != NNAB_FMT_PHASE_UNIT)
    return NNAB_EINVAL;
  if ((rc = check_arch())) return rc;
  FramedProblem p{};
  p.x = x; p.B = B; p.L = L; p.x_pitch = x_pitch;
  p.w_re = k_real; p.w_im = k_imag; p.F = n_bins; p.K = width; p.hop = hop;
  p.pad = pad; p.pad_mode = pad_mode; p.scale = scale; p.scale_all = scale_all;
  p.fmt = out_format; p.eps = sqrt_eps; p.power = 1.f; p.out = out; p.T = T;
  p.out_bins = n_bins; p.bin_offset = 0;
  p.h_k_begin = h_k_begin; p.h_k_end = h_k_end;
  return run_framed(p, packed, workspace, ws_bytes, path, (cudaStream_t)stream);
}

// ------------------------------------------------ CQT2010v2 / VQT pyramid ----
// Level lengths follow conv1d(stride=n, padding=127, kernel=256): (len - 2)/n + 1.
static inline int64_t decimated_len(int64_t len, int factor) {
  return len < 2 ? 0 : (len - 2) / factor + 1;
}

static size_t pyramid_level_bytes(int64_t B, int64_t L, int early_factor) {
  // [early (B, L0)] + ping/pong level buffers (B, <= L0/2 + 1)
  const int64_t L0 = early_factor > 1 ? decimated_len(L, early_factor) : L;
  size_t n = 0;
  if (early_factor > 1) n += align_up((size_t)B * align_up((size_t)L0, 4) * sizeof(float), 256);
  const size_t half = align_up((size_t)(L0 / 2 + 1), 4);
  n += 2 * align_up((size_t)B * half * sizeof(float), 256);
  return n;
}

size_t nnab_cqt_pyramid_workspace_bytes(int64_t B, int64_t L, int n_octaves, int early_factor,
                                        int max_width, int hop, int path) {
  (void)n_octaves;
  size_t n = pyramid_level_bytes(B, L, early_factor);
  if (path != NNAB_PATH_SIMT) {
    // split-signal scratch of the largest level (level 0), reused by every octave
    const int64_t L0 = early_factor > 1 ? decimated_len(L, early_factor) : L;
    n += tc_workspace_bytes(B, L0, max_width, hop, max_width / 2);
  }
  return n;
}

int nnab_cqt_pyramid_forward(const float* x, int64_t B, int64_t L, int64_t x_pitch, int n_octaves,
                             const float* const* h_k_real, const float* const* h_k_imag,
                             const void* const* h_packed, const int32_t* h_widths, int n_filters,
                             const float* lowpass,
                             const float* early_filter, int early_factor, int hop, int pad_mode,
                             int n_bins, const float* scale, float scale_all, int out_format,
                             float sqrt_eps, float* out, int64_t T, void* workspace,
                             size_t ws_bytes, int path, void* stream) {
  if (x == nullptr || out == nullptr || h_k_real == nullptr || h_k_imag == nullptr ||
      h_widths == nullptr || lowpass == nullptr || B < 0 || L <= 0 || x_pitch < L ||
      n_octaves <= 0 || n_filters <= 0 || hop <= 0 || n_bins <= 0 || early_factor < 1)
    return NNAB_EINVAL;
  if (early_factor > 1 && early_filter == nullptr) return NNAB_EINVAL;
  if (out_format != NNAB_FMT_MAGNITUDE && out_format != NNAB_FMT_COMPLEX &&
      out_format != NNAB_FMT_PHASE_UNIT)
    return NNAB_EINVAL;
  if (pad_mode != NNAB_PAD_REFLECT && pad_mode != NNAB_PAD_CONSTANT) return NNAB_EINVAL;
  int rc = check_arch();
  if (rc) return rc;
  int max_width = 0;
  for (int i = 0; i < n_octaves; ++i) max_width = h_widths[i] > max_width ? h_widths[i] : max_width;
  const size_t need =
      nnab_cqt_pyramid_workspace_bytes(B, L, n_octaves, early_factor, max_width, hop, path);
  if (need > 0 && (workspace == nullptr || ws_bytes < need)) return NNAB_EWORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;

  const size_t level_bytes = pyramid_level_bytes(B, L, early_factor);
  char* tc_ws = (char*)workspace + level_bytes;
  const size_t tc_ws_bytes = ws_bytes - level_bytes;
  char* wsp = (char*)workspace;
  const float* cur = x;
  int64_t cur_len = L, cur_pitch = x_pitch;
  if (early_factor > 1) {
    const int64_t L0 = decimated_len(L, early_factor);
    const int64_t pitch0 = (int64_t)align_up((size_t)L0, 4);
    float* e = (float*)wsp;
    wsp += align_up((size_t)B * pitch0 * sizeof(float), 256);
    if ((rc = launch_fir_decimate(x, B, L, x_pitch, early_filter, 256, early_factor, e, L0,
                                  pitch0, s)))
      return rc;
    cur = e; cur_len = L0; cur_pitch = pitch0;
  }
  const int64_t half_pitch = (int64_t)align_up((size_t)(cur_len / 2 + 1), 4);
  float* pingpong[2];
  pingpong[0] = (float*)wsp;
  pingpong[1] = (float*)(wsp + align_up((size_t)B * half_pitch * sizeof(float), 256));

  int cur_hop = hop;
  for (int i = 0; i < n_octaves; ++i) {
    if (i > 0) {
      const int64_t nl = decimated_len(cur_len, 2);
      float* dst = pingpong[i & 1];
      if ((rc = launch_fir_decimate(cur, B, cur_len, cur_pitch, lowpass, 256, 2, dst, nl,
                                    half_pitch, s)))
        return rc;
      cur = dst; cur_len = nl; cur_pitch = half_pitch;
      cur_hop /= 2;
    }
    if (cur_hop <= 0 || cur_len <= 0) return NNAB_EINVAL;
    const int width = h_widths[i];
    const int pad = width / 2;
    // get_cqt_complex: reflect padding that torch would reject falls back to zero padding.
    int mode = pad_mode;
    if (mode == NNAB_PAD_REFLECT && pad >= cur_len) mode = NNAB_PAD_CONSTANT;
    if (frames_of(cur_len, width, cur_hop, pad) != T) return NNAB_EINVAL;
    FramedProblem p{};
    p.x = cur; p.B = B; p.L = cur_len; p.x_pitch = cur_pitch;
    p.w_re = h_k_real[i]; p.w_im = h_k_imag[i]; p.F = n_filters; p.K = width; p.hop = cur_hop;
    p.pad = pad; p.pad_mode = mode; p.scale = nullptr; p.scale_all = scale_all;
    p.fmt = out_format; p.eps = sqrt_eps; p.power = 1.f; p.out = out; p.T = T;
    p.out_bins = n_bins;
    // octave i (0 = top) lands n_filters*(i+1) rows below the top of the output
    p.bin_offset = n_bins - n_filters * (i + 1);
    // per-bin scale is indexed by OUTPUT row: shift the pointer by the same offset
    p.scale = scale ? scale + p.bin_offset : nullptr;
    const void* pk = (h_packed != nullptr) ? h_packed[i] : nullptr;
    if (pk != nullptr && path != NNAB_PATH_SIMT && tc_supported(p) &&
        tc_ws_bytes >= tc_workspace_bytes(B, cur_len, width, cur_hop, pad)) {
      if ((rc = run_framed(p, pk, tc_ws, tc_ws_bytes, NNAB_PATH_TCGEN05, s))) return rc;
    } else {
      if (path == NNAB_PATH_TCGEN05) return NNAB_EALIGN;
      if ((rc = run_framed(p, nullptr, nullptr, 0, NNAB_PATH_SIMT, s))) return rc;
    }
  }
  return NNAB_OK;
}

}  // extern "C"
