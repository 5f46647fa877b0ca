// tezip_b200 -- PredNet inference handle: parameter upload, the input-independent t=0 maps, and the fp32
// CUDA-core ("direct") kernels.  The direct kernels compute the t=0 maps at create time for both paths and
// serve as the fp32 validation path (TZ_PREDNET_FP32_DIRECT); the product path is tz_conv_tc.cu.
//
// What `next(A)` = Model.predict([A, 0])[0,1] really needs (prednet.py:235-308):
//   t=0, zero state:  R0_l, C0_l, Ahat0_l do not depend on A            -> computed once in create()
//                     e_0 = [relu(P0 - A), relu(A - P0)],  a_{l+1} = pool(relu(conv_a(e_l))),
//                     e_{l+1} = [relu(Ahat0_{l+1} - a_{l+1}), relu(a_{l+1} - Ahat0_{l+1})]
//   t=1:              top-down gates on [R0_l | e_l | up(r_{l+1})]; the R0_l slice of the kernel is folded
//                     into the per-pixel bias map BM_l; c_l = f*C0_l + i*tanh(.), r_l = o*tanh(c_l)
//                     prediction = min(relu(conv_ahat0(r_0)), pixel_max)
//   (the t=1 input frame of zeros, and the t=1 A/Ahat/E units above layer 0, never reach the output.)
#include "tz_prednet.cuh"

#include <string.h>

namespace tz {

void *dev_alloc(tz_prednet *h, size_t bytes) {
  void *p = nullptr;
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    cudaGetLastError();
    return nullptr;
  }
  h->allocs.push_back(p);
  h->dev_bytes += (long long)bytes;
  return p;
}

// ------------------------------------------------------------------------------------------------ direct conv
constexpr int PX = 4;  // output pixels along x per thread (kernel-weight reuse in registers)

__global__ void __launch_bounds__(128) conv3x3_direct_kernel(ConvSrc s0, ConvSrc s1, int nsrc,
                                                             const float *__restrict__ Wt, int cin_w, int cout,
                                                             const float *__restrict__ bias,
                                                             const float *__restrict__ biasmap,
                                                             float *__restrict__ out, int B, int H, int W, int act,
                                                             float clip) {
  const int xg_n = (W + PX - 1) / PX;
  long long total = (long long)B * H * xg_n * cout;
  long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= total) return;
  int co = (int)(id % cout);
  long long t = id / cout;
  int xg = (int)(t % xg_n);
  t /= xg_n;
  int y = (int)(t % H);
  int b = (int)(t / H);
  const int x0 = xg * PX;
  float acc[PX];
#pragma unroll
  for (int p = 0; p < PX; p++) acc[p] = 0.0f;
  for (int ky = 0; ky < 3; ky++) {
    int yy = y + ky - 1;
    if (yy < 0 || yy >= H) continue;
    for (int kx = 0; kx < 3; kx++) {
      for (int si = 0; si < nsrc; si++) {
        const ConvSrc &s = si == 0 ? s0 : s1;
        const int Hs = s.up ? (H >> 1) : H, Ws = s.up ? (W >> 1) : W;
        const float *wp = Wt + ((long long)(ky * 3 + kx) * cin_w + s.wofs) * cout + co;
        const float *base = s.ptr + (long long)b * s.bstride + (long long)(s.up ? (yy >> 1) : yy) * Ws * s.C;
        (void)Hs;
        const float *ip[PX];
        bool ok[PX];
#pragma unroll
        for (int p = 0; p < PX; p++) {
          int xx = x0 + p + kx - 1;
          ok[p] = (xx >= 0 && xx < W && x0 + p < W);
          ip[p] = base + (long long)(ok[p] ? (s.up ? (xx >> 1) : xx) : 0) * s.C;
        }
        for (int ci = 0; ci < s.C; ci++) {
          float w = wp[(long long)ci * cout];
#pragma unroll
          for (int p = 0; p < PX; p++)
            if (ok[p]) acc[p] = fmaf(ip[p][ci], w, acc[p]);
        }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < PX; p++) {
    int x = x0 + p;
    if (x < W) {
      float v = acc[p] + (biasmap ? biasmap[((long long)y * W + x) * cout + co] : bias[co]);
      if (act >= 1) v = fmaxf(v, 0.0f);
      if (act == 2) v = fminf(v, clip);
      out[(((long long)b * H + y) * W + x) * cout + co] = v;
    }
  }
}

int conv3x3_direct(const ConvSrc *srcs, int nsrc, const float *Wt, int cin_w, int cout, const float *bias,
                   const float *biasmap, float *out, int B, int H, int W, int act, float clip,
                   cudaStream_t st) {
  ConvSrc z = {nullptr, 0, 0, 0, 0};
  ConvSrc s0 = nsrc > 0 ? srcs[0] : z, s1 = nsrc > 1 ? srcs[1] : z;
  long long total = (long long)B * H * ((W + PX - 1) / PX) * cout;
  if (total == 0) return TZ_OK;
  long long blocks = (total + 127) / 128;
  conv3x3_direct_kernel<<<(unsigned)blocks, 128, 0, st>>>(s0, s1, nsrc, Wt, cin_w, cout, bias, biasmap, out, B, H, W,
                                                         act, clip);
  TZ_CHECK_LAUNCH();
  return TZ_OK;
}

// keras hard_sigmoid: clip(0.2*x + 0.5, 0, 1), two rounded operations as TF evaluates it
__device__ __forceinline__ float hard_sigmoid(float x) {
  return fminf(fmaxf(__fadd_rn(__fmul_rn(0.2f, x), 0.5f), 0.0f), 1.0f);
}

// prednet.py:255-259.  pre [B,H,W,4R] with blocks i,f,c,o; c_prev [H,W,R] broadcast over the batch or null (= 0).
__global__ void __launch_bounds__(256) lstm_direct_kernel(const float *__restrict__ pre,
                                                          const float *__restrict__ c_prev,
                                                          float *__restrict__ r_out, float *__restrict__ c_out,
                                                          long long total, int HW, int R) {
  long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= total) return;
  int ch = (int)(id % R);
  long long pix = id / R;  // b*HW + p
  const float *p = pre + pix * 4 * R;
  float i = hard_sigmoid(p[ch]);
  float f = hard_sigmoid(p[R + ch]);
  float g = tanhf(p[2 * R + ch]);
  float o = hard_sigmoid(p[3 * R + ch]);
  float cp = c_prev ? c_prev[(pix % HW) * R + ch] : 0.0f;
  float c = __fadd_rn(__fmul_rn(f, cp), __fmul_rn(i, g));
  float r = __fmul_rn(o, tanhf(c));
  r_out[id] = r;
  if (c_out) c_out[id] = c;
}

int lstm_direct(const float *pre, const float *c_prev, float *r_out, float *c_out, int B, int H, int W, int R,
                cudaStream_t st) {
  long long total = (long long)B * H * W * R;
  if (total == 0) return TZ_OK;
  lstm_direct_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pre, c_prev, r_out, c_out, total, H * W, R);
  TZ_CHECK_LAUNCH();
  return TZ_OK;
}

// prednet.py:274-277 at layer 0, t=0: e_0 = [relu(P0 - A), relu(A - P0)]
__global__ void __launch_bounds__(256) e0_direct_kernel(const float *__restrict__ in, const float *__restrict__ p0,
                                                        float *__restrict__ e, long long total, long long frame,
                                                        int C) {
  long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= total) return;
  int c = (int)(id % C);
  long long pix = id / C;
  float a = in[id];
  float ah = p0[id % frame];
  e[pix * 2 * C + c] = fmaxf(__fsub_rn(ah, a), 0.0f);
  e[pix * 2 * C + C + c] = fmaxf(__fsub_rn(a, ah), 0.0f);
}

// prednet.py:290-291 then :274-277 one layer up: a = maxpool2(relu(conv + bias)) (the conv kernel already
// applied bias + relu); e = [relu(Ahat0 - a), relu(a - Ahat0)].
__global__ void __launch_bounds__(256) pool_e_direct_kernel(const float *__restrict__ a_full,
                                                            const float *__restrict__ ahat0,
                                                            float *__restrict__ e, long long total, int Ho, int Wo,
                                                            int S) {
  long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= total) return;
  int c = (int)(id % S);
  long long t = id / S;
  int x = (int)(t % Wo);
  t /= Wo;
  int y = (int)(t % Ho);
  long long b = t / Ho;
  const int Wi = Wo * 2;
  const float *p = a_full + (((b * (Ho * 2) + 2 * y) * Wi) + 2 * x) * S + c;
  float a = fmaxf(fmaxf(p[0], p[S]), fmaxf(p[(long long)Wi * S], p[(long long)Wi * S + S]));
  float ah = ahat0[((long long)y * Wo + x) * S + c];
  long long pix = (b * Ho + y) * Wo + x;
  e[pix * 2 * S + c] = fmaxf(__fsub_rn(ah, a), 0.0f);
  e[pix * 2 * S + S + c] = fmaxf(__fsub_rn(a, ah), 0.0f);
}

static int direct_next_chunk(tz_prednet *h, const float *in, float *out, int B, cudaStream_t st) {
  const int L = h->L;
  const int C = h->S[0];
  {
    long long total = (long long)B * h->H[0] * h->W[0] * C;
    e0_direct_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, h->Ahat0[0], h->e[0], total,
                                                                     (long long)h->H[0] * h->W[0] * C, C);
    TZ_CHECK_LAUNCH();
  }
  for (int l = 0; l < L - 1; l++) {  // bottom-up targets/errors at t=0
    ConvSrc s = {h->e[l], 2 * h->S[l], 0, 0, (long long)h->H[l] * h->W[l] * 2 * h->S[l]};
    int rc = conv3x3_direct(&s, 1, h->w_a[l], 2 * h->S[l], h->S[l + 1], h->b_a[l], nullptr, h->pre, B, h->H[l],
                            h->W[l], 1, 0.0f, st);
    if (rc) return rc;
    long long total = (long long)B * h->H[l + 1] * h->W[l + 1] * h->S[l + 1];
    pool_e_direct_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(h->pre, h->Ahat0[l + 1], h->e[l + 1], total,
                                                                         h->H[l + 1], h->W[l + 1], h->S[l + 1]);
    TZ_CHECK_LAUNCH();
  }
  for (int l = L - 1; l >= 0; l--) {  // top-down representation update at t=1
    ConvSrc s[2];
    int ns = 0;
    s[ns++] = {h->e[l], 2 * h->S[l], h->R[l], 0, (long long)h->H[l] * h->W[l] * 2 * h->S[l]};
    if (l < L - 1)
      s[ns++] = {h->r[l + 1], h->R[l + 1], h->R[l] + 2 * h->S[l], 1,
                 (long long)h->H[l + 1] * h->W[l + 1] * h->R[l + 1]};
    int rc = conv3x3_direct(s, ns, h->w_g[l], h->cin_g[l], 4 * h->R[l], nullptr, h->BM[l], h->pre, B, h->H[l],
                            h->W[l], 0, 0.0f, st);
    if (rc) return rc;
    rc = lstm_direct(h->pre, h->C0[l], h->r[l], nullptr, B, h->H[l], h->W[l], h->R[l], st);
    if (rc) return rc;
  }
  ConvSrc s = {h->r[0], h->R[0], 0, 0, (long long)h->H[0] * h->W[0] * h->R[0]};
  return conv3x3_direct(&s, 1, h->w_ahat[0], h->R[0], h->S[0], h->b_ahat[0], nullptr, out, B, h->H[0], h->W[0], 2,
                        h->cfg.pixel_max, st);
}

}  // namespace tz

using namespace tz;

// index of (key, layer) in the reference's weight list (prednet.py:212-227)
static int widx(int L, int key /*0 a,1 ahat,2 c,3 f,4 i,5 o*/, int l) {
  int n = 0;
  for (int k = 0; k < key; k++) n += (k == 0) ? (L - 1) : L;
  return 2 * (n + l);
}

static int init_constants(tz_prednet *h) {
  cudaStream_t st = 0;
  const int L = h->L;
  float *pre = nullptr;
  size_t mx = 0;
  for (int l = 0; l < L; l++) {
    size_t v = (size_t)h->H[l] * h->W[l] * 4 * h->R[l];
    if (v > mx) mx = v;
  }
  TZ_CHECK_CUDA(cudaMalloc(&pre, mx * sizeof(float)));
  int rc = TZ_OK;
  for (int l = L - 1; l >= 0 && rc == TZ_OK; l--) {  // prednet.py:249-264 at t=0: inputs [0 | 0 | up(R0_{l+1})]
    ConvSrc s = {nullptr, 0, 0, 0, 0};
    int ns = 0;
    if (l < L - 1) {
      s = {h->R0[l + 1], h->R[l + 1], h->R[l] + 2 * h->S[l], 1, 0};
      ns = 1;
    }
    rc = conv3x3_direct(&s, ns, h->w_g[l], h->cin_g[l], 4 * h->R[l], h->b_g[l], nullptr, pre, 1, h->H[l], h->W[l], 0,
                        0.0f, st);
    if (rc == TZ_OK) rc = lstm_direct(pre, nullptr, h->R0[l], h->C0[l], 1, h->H[l], h->W[l], h->R[l], st);
  }
  for (int l = 0; l < L && rc == TZ_OK; l++) {
    ConvSrc s = {h->R0[l], h->R[l], 0, 0, 0};
    // prednet.py:268-270: Ahat0_l = relu(conv(R0_l)); layer 0 is clipped to pixel_max (= P0)
    rc = conv3x3_direct(&s, 1, h->w_ahat[l], h->R[l], h->S[l], h->b_ahat[l], nullptr, h->Ahat0[l], 1, h->H[l],
                        h->W[l], l == 0 ? 2 : 1, h->cfg.pixel_max, st);
    // t=1 hoist: BM_l = b_g + conv(R0_l, W_g[:, :, 0:R_l, :])
    if (rc == TZ_OK)
      rc = conv3x3_direct(&s, 1, h->w_g[l], h->cin_g[l], 4 * h->R[l], h->b_g[l], nullptr, h->BM[l], 1, h->H[l],
                          h->W[l], 0, 0.0f, st);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(pre);
  if (rc != TZ_OK) return rc;
  if (e != cudaSuccess) {
    set_error("PredNet constant initialisation failed: %s", cudaGetErrorString(e));
    return TZ_ECUDA;
  }
  return TZ_OK;
}

extern "C" {

int tz_prednet_create(const tz_prednet_config *cfg, const float *const *weights_host,
                      const long long *weight_elems, int n_weights, tz_prednet **out) {
  TZ_REQUIRE(cfg && weights_host && weight_elems && out, "tz_prednet_create: null argument");
  const int L = cfg->n_layers;
  TZ_REQUIRE(L >= 2 && L <= TZ_MAX_LAYERS, "tz_prednet_create: n_layers %d out of range", L);
  TZ_REQUIRE(n_weights == 2 * (6 * L - 1), "tz_prednet_create: expected %d weight arrays, got %d", 2 * (6 * L - 1),
             n_weights);
  TZ_REQUIRE(cfg->Hp > 0 && cfg->Wp > 0 && cfg->Hp % (1 << (L - 1)) == 0 && cfg->Wp % (1 << (L - 1)) == 0,
             "tz_prednet_create: Hp x Wp = %d x %d must be multiples of %d", cfg->Hp, cfg->Wp, 1 << (L - 1));
  TZ_REQUIRE(cfg->max_batch >= 1, "tz_prednet_create: max_batch must be >= 1");
  for (int l = 0; l < L; l++)
    TZ_REQUIRE(cfg->stack_sizes[l] > 0 && cfg->r_stack_sizes[l] > 0, "tz_prednet_create: bad channel count");
  int ndev = 0;
  TZ_CHECK_CUDA(cudaGetDeviceCount(&ndev));
  TZ_REQUIRE(cfg->device >= 0 && cfg->device < ndev, "tz_prednet_create: no CUDA device %d", cfg->device);
  int major = 0;
  TZ_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, cfg->device));
  if (major != 10) {
    set_error("tz_prednet_create: device %d has compute capability %d.x; this library is built for sm_100a only",
              cfg->device, major);
    return TZ_ECUDA;
  }
  TZ_CHECK_CUDA(cudaSetDevice(cfg->device));

  tz_prednet *h = new tz_prednet();
  memset(&h->cfg, 0, sizeof(h->cfg));
  h->cfg = *cfg;
  h->L = L;
  h->device = cfg->device;
  h->direct = (cfg->flags & TZ_PREDNET_FP32_DIRECT) != 0;
  h->dev_bytes = 0;
  h->tc = nullptr;
  for (int l = 0; l < L; l++) {
    h->S[l] = cfg->stack_sizes[l];
    h->R[l] = cfg->r_stack_sizes[l];
    h->H[l] = cfg->Hp >> l;
    h->W[l] = cfg->Wp >> l;
  }
  for (int l = 0; l < L; l++) h->cin_g[l] = 2 * h->S[l] + h->R[l] + (l < L - 1 ? h->R[l + 1] : 0);

  // ---- validate sizes, repack gates to [3,3,Cin,4R] (blocks i,f,c,o), upload
  std::vector<std::vector<float>> packed;  // handed to the tensor-core path as well
  int rc = TZ_OK;
  auto fail = [&](int code) {
    tz_prednet_destroy(h);
    return code;
  };
  auto upload = [&](const float *src, size_t n) -> float * {
    float *d = (float *)dev_alloc(h, n * sizeof(float));
    if (!d) return nullptr;
    if (cudaMemcpy(d, src, n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
      set_error("weight upload failed");
      return nullptr;
    }
    return d;
  };
  for (int l = 0; l < L; l++) {
    const int gate_key[4] = {4, 3, 2, 5};  // i, f, c, o
    long long kn = 9LL * h->cin_g[l] * h->R[l];
    std::vector<float> wg((size_t)kn * 4), bg((size_t)4 * h->R[l]);
    for (int g = 0; g < 4; g++) {
      int wi = widx(L, gate_key[g], l);
      if (weight_elems[wi] != kn || weight_elems[wi + 1] != h->R[l]) {
        set_error("tz_prednet_create: gate %d layer %d: expected %lld/%d elements, got %lld/%lld", g, l, kn, h->R[l],
                  weight_elems[wi], weight_elems[wi + 1]);
        return fail(TZ_EINVAL);
      }
      const float *src = weights_host[wi];
      for (long long k = 0; k < 9LL * h->cin_g[l]; k++)
        for (int co = 0; co < h->R[l]; co++) wg[(size_t)k * 4 * h->R[l] + g * h->R[l] + co] = src[k * h->R[l] + co];
      for (int co = 0; co < h->R[l]; co++) bg[(size_t)g * h->R[l] + co] = weights_host[wi + 1][co];
    }
    h->w_g[l] = upload(wg.data(), wg.size());
    h->b_g[l] = upload(bg.data(), bg.size());
    if (!h->w_g[l] || !h->b_g[l]) return fail(TZ_ECUDA);
    int wi = widx(L, 1, l);
    long long an = 9LL * h->R[l] * h->S[l];
    if (weight_elems[wi] != an || weight_elems[wi + 1] != h->S[l]) {
      set_error("tz_prednet_create: ahat layer %d: wrong element count", l);
      return fail(TZ_EINVAL);
    }
    h->w_ahat[l] = upload(weights_host[wi], (size_t)an);
    h->b_ahat[l] = upload(weights_host[wi + 1], (size_t)h->S[l]);
    if (!h->w_ahat[l] || !h->b_ahat[l]) return fail(TZ_ECUDA);
    if (l < L - 1) {
      wi = widx(L, 0, l);
      long long n = 9LL * 2 * h->S[l] * h->S[l + 1];
      if (weight_elems[wi] != n || weight_elems[wi + 1] != h->S[l + 1]) {
        set_error("tz_prednet_create: a layer %d: wrong element count", l);
        return fail(TZ_EINVAL);
      }
      h->w_a[l] = upload(weights_host[wi], (size_t)n);
      h->b_a[l] = upload(weights_host[wi + 1], (size_t)h->S[l + 1]);
      if (!h->w_a[l] || !h->b_a[l]) return fail(TZ_ECUDA);
    }
    packed.push_back(std::move(wg));
  }

  // ---- constants
  for (int l = 0; l < L; l++) {
    size_t hw = (size_t)h->H[l] * h->W[l];
    h->R0[l] = (float *)dev_alloc(h, hw * h->R[l] * sizeof(float));
    h->C0[l] = (float *)dev_alloc(h, hw * h->R[l] * sizeof(float));
    h->Ahat0[l] = (float *)dev_alloc(h, hw * h->S[l] * sizeof(float));
    h->BM[l] = (float *)dev_alloc(h, hw * 4 * h->R[l] * sizeof(float));
    if (!h->R0[l] || !h->C0[l] || !h->Ahat0[l] || !h->BM[l]) return fail(TZ_ENOMEM);
  }
  rc = init_constants(h);
  if (rc != TZ_OK) return fail(rc);

  if (h->direct) {
    h->direct_chunk = cfg->max_batch < 8 ? cfg->max_batch : 8;
    size_t mx = 0;
    for (int l = 0; l < L; l++) {
      size_t hw = (size_t)h->H[l] * h->W[l];
      size_t v = hw * 4 * h->R[l];
      if (l < L - 1 && hw * h->S[l + 1] > v) v = hw * h->S[l + 1];
      if (v > mx) mx = v;
      h->e[l] = (float *)dev_alloc(h, (size_t)h->direct_chunk * hw * 2 * h->S[l] * sizeof(float));
      h->r[l] = (float *)dev_alloc(h, (size_t)h->direct_chunk * hw * h->R[l] * sizeof(float));
      if (!h->e[l] || !h->r[l]) return fail(TZ_ENOMEM);
    }
    h->pre = (float *)dev_alloc(h, (size_t)h->direct_chunk * mx * sizeof(float));
    if (!h->pre) return fail(TZ_ENOMEM);
  } else {
    rc = tc_create(h, packed);
    if (rc != TZ_OK) return fail(rc);
  }
  *out = h;
  return TZ_OK;
}

int tz_prednet_destroy(tz_prednet *h) {
  if (!h) return TZ_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  if (h->tc) tc_destroy(h);
  for (void *p : h->allocs) cudaFree(p);
  delete h;
  return TZ_OK;
}

int tz_prednet_p0(tz_prednet *h, float *out, void *stream) {
  TZ_REQUIRE(h && out, "tz_prednet_p0: null argument");
  TZ_CHECK_CUDA(cudaMemcpyAsync(out, h->Ahat0[0], (size_t)h->H[0] * h->W[0] * h->S[0] * sizeof(float),
                                cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return TZ_OK;
}

int tz_prednet_next(tz_prednet *h, const float *in, float *out, int B, void *stream) {
  TZ_REQUIRE(h && in && out, "tz_prednet_next: null argument");
  TZ_REQUIRE(B >= 0 && B <= h->cfg.max_batch, "tz_prednet_next: B=%d exceeds max_batch=%d", B, h->cfg.max_batch);
  if (B == 0) return TZ_OK;
  cudaStream_t st = (cudaStream_t)stream;
  h->last_out = nullptr;
  if (!h->direct) {
    int rc = tc_next(h, in, out, B, st);
    if (rc == TZ_OK) {
      h->last_out = out;
      h->last_B = B;
    }
    return rc;
  }
  const long long frame = (long long)h->H[0] * h->W[0] * h->S[0];
  for (int b0 = 0; b0 < B; b0 += h->direct_chunk) {
    int nb = B - b0 < h->direct_chunk ? B - b0 : h->direct_chunk;
    int rc = direct_next_chunk(h, in + b0 * frame, out + b0 * frame, nb, st);
    if (rc) return rc;
  }
  h->last_out = out;
  h->last_B = B;
  return TZ_OK;
}

int tz_prednet_next_chained(tz_prednet *h, float *out, int B, void *stream) {
  TZ_REQUIRE(h && out, "tz_prednet_next_chained: null argument");
  TZ_REQUIRE(h->last_out != nullptr, "tz_prednet_next_chained: no previous tz_prednet_next on this handle");
  TZ_REQUIRE(B >= 0 && B <= h->last_B, "tz_prednet_next_chained: B=%d exceeds the previous step's %d frames", B,
             h->last_B);
  TZ_REQUIRE(out != h->last_out, "tz_prednet_next_chained: out must not be the previous prediction");
  if (B == 0) return TZ_OK;
  const float *in = h->last_out;
  if (h->direct || !h->x0_staged) return tz_prednet_next(h, in, out, B, stream);
  int rc = tc_next(h, in, out, B, (cudaStream_t)stream, nullptr, true);
  h->last_out = rc == TZ_OK ? out : nullptr;
  h->last_B = B;
  return rc;
}

int tz_prednet_kernel_count(tz_prednet *h) {
  if (!h) return 0;
  return h->direct ? 0 : 2 * h->L + 1;   // e0, a_0..a_{L-2}, gates_{L-1}..gates_0, ahat_0
}

int tz_prednet_kernel_info(tz_prednet *h, int i, char *name, int name_len, double *flops_per_frame) {
  TZ_REQUIRE(h && name && flops_per_frame && name_len > 0, "tz_prednet_kernel_info: null argument");
  const int L = h->L, n = 2 * L + 1;
  TZ_REQUIRE(!h->direct && i >= 0 && i < n, "tz_prednet_kernel_info: no kernel %d", i);
  double fl = 0.0;
  if (i == 0) {
    snprintf(name, name_len, "e0");
  } else if (i <= L - 1) {
    int l = i - 1;
    snprintf(name, name_len, "conv_tc_a%d", l);
    fl = 2.0 * h->H[l] * h->W[l] * 9.0 * 2 * h->S[l] * h->S[l + 1];
  } else if (i <= 2 * L - 1) {
    int l = L - 1 - (i - L);
    fl = 2.0 * h->H[l] * h->W[l] * 9.0 * h->cin_g[l] * 4 * h->R[l];
    if (l == 0 && tc_layer0_split(h)) {
      // layer 0 is split: its up(r_1) half runs on the tensor core at r_1's resolution (this launch); the e_0 half,
      // the LSTM, A-hat_0 and the next step's error units are the "l0_tail" launch
      snprintf(name, name_len, "conv_tc_gates0_r1half");
      fl = 2.0 * h->H[0] * h->W[0] * 9.0 * h->R[1] * 4 * h->R[0];
    } else {
      snprintf(name, name_len, "conv_tc_gates%d", l);
    }
  } else {
    fl = 2.0 * h->H[0] * h->W[0] * 9.0 * h->R[0] * h->S[0];
    if (tc_layer0_split(h)) {
      snprintf(name, name_len, "l0_tail");
      fl += 2.0 * h->H[0] * h->W[0] * 9.0 * (h->cin_g[0] - h->R[1]) * 4 * h->R[0];
    } else {
      snprintf(name, name_len, "ahat0");
    }
  }
  *flops_per_frame = fl;
  return TZ_OK;
}

int tz_prednet_next_timed(tz_prednet *h, const float *in, float *out, int B, void *stream, float *ms, int n_ms) {
  TZ_REQUIRE(h && in && out && ms, "tz_prednet_next_timed: null argument");
  TZ_REQUIRE(!h->direct, "tz_prednet_next_timed: tensor-core path only");
  TZ_REQUIRE(B >= 1 && B <= h->cfg.max_batch, "tz_prednet_next_timed: bad B");
  const int n = 2 * h->L + 1;
  TZ_REQUIRE(n_ms >= n, "tz_prednet_next_timed: need room for %d timings", n);
  cudaEvent_t ev[2 * TZ_MAX_LAYERS + 2];
  for (int i = 0; i <= n; i++) TZ_CHECK_CUDA(cudaEventCreate(&ev[i]));
  h->last_out = nullptr;   // a timed step does not take part in chaining
  int rc = tc_next(h, in, out, B, (cudaStream_t)stream, ev);
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  if (rc == TZ_OK && e == cudaSuccess)
    for (int i = 0; i < n; i++) cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]);
  for (int i = 0; i <= n; i++) cudaEventDestroy(ev[i]);
  if (e != cudaSuccess) {
    set_error("tz_prednet_next_timed: %s", cudaGetErrorString(e));
    return TZ_ECUDA;
  }
  return rc;
}

long long tz_prednet_device_bytes(tz_prednet *h) { return h ? h->dev_bytes : 0; }

double tz_prednet_flops_per_frame(tz_prednet *h) {
  if (!h) return 0.0;
  double macs = 0.0;
  for (int l = 0; l < h->L; l++) {
    double hw = (double)h->H[l] * h->W[l];
    if (l < h->L - 1) macs += hw * 9.0 * 2 * h->S[l] * h->S[l + 1];  // a_l at t=0
    macs += hw * 9.0 * h->cin_g[l] * 4 * h->R[l];                    // i,f,c,o at t=1, full concatenated K
  }
  macs += (double)h->H[0] * h->W[0] * 9.0 * h->R[0] * h->S[0];       // ahat_0 at t=1
  return 2.0 * macs;
}

}  // extern "C"
