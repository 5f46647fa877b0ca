// tezip_b200 -- PredNet handle shared between the fp32 direct kernels (tz_prednet.cu) and the tcgen05
// implicit-GEMM path (tz_conv_tc.cu).
#pragma once
#include "tz_common.cuh"
#include <cuda_fp16.h>
#include <vector>

struct TcState;  // tensor-core path state (tz_conv_tc.cu)

struct tz_prednet {
  tz_prednet_config cfg;
  int L;
  int S[TZ_MAX_LAYERS], R[TZ_MAX_LAYERS], H[TZ_MAX_LAYERS], W[TZ_MAX_LAYERS];
  int cin_g[TZ_MAX_LAYERS];  // gate conv input channels, prednet.py:220-222
  int device;
  bool direct;
  int direct_chunk;  // frames per pass of the fp32 direct path
  long long dev_bytes;
  std::vector<void *> allocs;

  // fp32 parameters on the device.  Gate kernels are packed [3,3,Cin,4R] with output blocks i,f,c,o.
  float *w_a[TZ_MAX_LAYERS], *b_a[TZ_MAX_LAYERS];
  float *w_ahat[TZ_MAX_LAYERS], *b_ahat[TZ_MAX_LAYERS];
  float *w_g[TZ_MAX_LAYERS], *b_g[TZ_MAX_LAYERS];

  // input-independent t=0 maps (prednet.py:249-271 with zero state), batch-broadcast
  float *R0[TZ_MAX_LAYERS];     // [H_l,W_l,R_l]   r after t=0
  float *C0[TZ_MAX_LAYERS];     // [H_l,W_l,R_l]   c after t=0
  float *Ahat0[TZ_MAX_LAYERS];  // [H_l,W_l,S_l]   A-hat at t=0 (layer 0: clipped = P0)
  float *BM[TZ_MAX_LAYERS];     // [H_l,W_l,4R_l]  gate bias + conv(R0_l, W_g[:, :, 0:R_l, :])  (t=1, hoisted)

  // fp32 direct-path workspaces
  float *e[TZ_MAX_LAYERS];  // [chunk,H_l,W_l,2S_l]
  float *r[TZ_MAX_LAYERS];  // [chunk,H_l,W_l,R_l]
  float *pre;               // [chunk, max pre-activation plane]

  TcState *tc;

  // chained stepping (tz_prednet_next_chained): the last prediction written and how many frames of it are staged
  const float *last_out;
  int last_B;
  bool x0_staged;   // tensor-core path: X_0 already holds the error units of last_out (written by the ahat0 kernel)
};

namespace tz {

void *dev_alloc(tz_prednet *h, size_t bytes);  // tracked cudaMalloc (nullptr on failure, error set)

struct ConvSrc {
  const float *ptr;   // [B or 1, Hs, Ws, C]
  int C;              // channels of this source
  int wofs;           // first input-channel row of the kernel that multiplies this source
  int up;             // 1: source is at half resolution and is read through nearest 2x upsampling
  long long bstride;  // elements between batch items (0 = broadcast)
};

// out[b,y,x,co] = act(bias + sum_{ky,kx,src,ci} in * W), fixed summation order (ky, kx, src, ci).
int conv3x3_direct(const ConvSrc *srcs, int nsrc, const float *Wt, int cin_w, int cout, const float *bias,
                   const float *biasmap, float *out, int B, int H, int W, int act, float clip,
                   cudaStream_t st);
int lstm_direct(const float *pre, const float *c_prev, float *r_out, float *c_out, int B, int H, int W, int R,
                cudaStream_t st);

// tensor-core path (tz_conv_tc.cu)
int tc_create(tz_prednet *h, const std::vector<std::vector<float>> &w_host);
void tc_destroy(tz_prednet *h);
bool tc_layer0_split(tz_prednet *h);   // layer 0 runs as conv(r_1) at r_1's resolution + the fused tail kernel
// skip_e0: X_0 already holds the error units of `in` (staged by the previous step's ahat0 kernel)
int tc_next(tz_prednet *h, const float *in, float *out, int B, cudaStream_t st, cudaEvent_t *ev = nullptr,
            bool skip_e0 = false);

}  // namespace tz
