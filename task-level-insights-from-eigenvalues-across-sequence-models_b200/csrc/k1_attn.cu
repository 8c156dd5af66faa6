// k1_attn.cu -- K1'': linear-attention normaliser nu_t (O(T) prefix form of the reference's (B,T,T,H) score tensor) and
// the causal linear-attention layer forward in recurrent form.
//
// Reference operators:
//   get_eig_att_linear  analysis/eval_eig.py:97-135   nu_t = sum_{s<=t} phi(q_t).phi(k_s), phi = elu+1   (scores materialised O(T^2))
//   SelfLinAttention.forward  models/attention.py:63-83    kv cumsum materialised as (B,T,H,d,dv)
//   SelfNormAttention.forward models/norm_attention.py:61-89
// Roofline: HBM for nu (2*d*4 bytes read per eigenvalue, 8 written); FMA-pipe for the layer forward (2*d*dv FMA per token).
#include "linattn.cuh"
#include <stdlib.h>
#include <string.h>

namespace eigb200 {

// ---- nu: CTA per (b,h); 8 warps split T into contiguous ranges; two passes (range sums, then the walk) -------------------
constexpr int NU_WARPS = 8;
constexpr int NU_DMAX = 8;            // channels per lane (d <= 256)

__global__ void __launch_bounds__(NU_WARPS * 32) linattn_nu_kernel(const float* __restrict__ q, const float* __restrict__ k, int64_t ld,
                                                                   int64_t T, int H, int d, double* __restrict__ nu) {
  __shared__ double part[NU_WARPS][NU_DMAX * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h = blockIdx.x, b = blockIdx.y;
  const int64_t per = (T + NU_WARPS - 1) / NU_WARPS;
  const int64_t t0 = min(T, warp * per), t1 = min(T, t0 + per);
  const float* qb = q + ((size_t)b * T) * ld + (size_t)h * d;
  const float* kb = k + ((size_t)b * T) * ld + (size_t)h * d;
  double S[NU_DMAX];
#pragma unroll
  for (int i = 0; i < NU_DMAX; ++i) S[i] = 0.0;
  for (int64_t t = t0; t < t1; ++t) {
#pragma unroll
    for (int i = 0; i < NU_DMAX; ++i) {
      const int c = lane + 32 * i;
      if (c < d) S[i] += (double)(elu_f(__ldg(kb + t * ld + c)) + 1.f);
    }
  }
#pragma unroll
  for (int i = 0; i < NU_DMAX; ++i) part[warp][lane + 32 * i] = S[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NU_DMAX; ++i) {
    double acc = 0.0;
    for (int w = 0; w < warp; ++w) acc += part[w][lane + 32 * i];
    S[i] = acc;                                                    // exclusive prefix: state before this warp's range
  }
  for (int64_t t = t0; t < t1; ++t) {
    double dot = 0.0;
#pragma unroll
    for (int i = 0; i < NU_DMAX; ++i) {
      const int c = lane + 32 * i;
      if (c < d) {
        S[i] += (double)(elu_f(__ldg(kb + t * ld + c)) + 1.f);
        dot += (double)(elu_f(__ldg(qb + t * ld + c)) + 1.f) * S[i];
      }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane == 0) nu[((size_t)b * T + t) * H + h] = dot;
  }
}

// ---- layer forward: CTA per (b,h); thread (g, j) owns RI rows of column j of the d x dv state ----------------------
constexpr int LA_THREADS = 256;
constexpr int LA_TC = 16;             // tokens per staging round


template <int RI>
__global__ void __launch_bounds__(LA_THREADS) linattn_forward_kernel(const LinAttnParams p) {
  extern __shared__ __align__(16) float sm[];
  const int d = p.d, dv = p.dv;
  const int groups = LA_THREADS / dv;
  float* qs = sm;                                  // [TC][d]
  float* ks = qs + LA_TC * d;                      // [TC][d]
  float* vs = ks + LA_TC * d;                      // [TC][dv]
  float* pn = vs + LA_TC * dv;                     // [TC][groups][dv] partial numerators
  float* pd = pn + LA_TC * groups * dv;            // [TC][groups]     partial denominators
  const int tid = threadIdx.x;
  const int g = tid / dv, j = tid - g * dv;
  const int i0 = g * RI;
  const int h = blockIdx.x, b = blockIdx.y;
  const size_t rowbase = (size_t)b * p.T;
  const float* qb = p.q + (size_t)h * d;
  const float* kb = p.k + (size_t)h * d;
  const float* vb = p.v + (size_t)h * dv;

  float kv[RI], ksum[RI];
#pragma unroll
  for (int i = 0; i < RI; ++i) { kv[i] = 0.f; ksum[i] = 0.f; }

  for (int64_t t0 = 0; t0 < p.T; t0 += LA_TC) {
    const int tc = (int)min((int64_t)LA_TC, p.T - t0);
    __syncthreads();
    for (int i = tid; i < tc * d; i += LA_THREADS) {
      const int r = i / d, c = i - r * d;
      float qv = __ldg(qb + (rowbase + t0 + r) * p.ld + c), kvv = __ldg(kb + (rowbase + t0 + r) * p.ld + c);
      if (p.phi_elu) { qv = elu_f(qv) + 1.f; kvv = elu_f(kvv) + 1.f; }
      qs[i] = qv; ks[i] = kvv;
    }
    for (int i = tid; i < tc * dv; i += LA_THREADS) {
      const int r = i / dv, c = i - r * dv;
      vs[i] = __ldg(vb + (rowbase + t0 + r) * p.ld + c);
    }
    __syncthreads();
    for (int tt = 0; tt < tc; ++tt) {
      const float vj = vs[tt * dv + j];
      float num = 0.f, den = 0.f;
#pragma unroll
      for (int i = 0; i < RI; ++i) {
        const float ki = ks[tt * d + i0 + i], qi = qs[tt * d + i0 + i];
        kv[i] = fmaf(ki * p.kscale, vj, kv[i]);
        num = fmaf(qi, kv[i], num);
        if (p.normalise) { ksum[i] += ki; den = fmaf(qi, ksum[i], den); }
      }
      pn[(tt * groups + g) * dv + j] = num;
      if (p.normalise && j == 0) pd[tt * groups + g] = den;
    }
    __syncthreads();
    for (int i = tid; i < tc * dv; i += LA_THREADS) {
      const int r = i / dv, c = i - r * dv;
      float num = 0.f;
      for (int gg = 0; gg < groups; ++gg) num += pn[(r * groups + gg) * dv + c];
      float scale = 1.f;
      if (p.normalise) {
        float den = 0.f;
        for (int gg = 0; gg < groups; ++gg) den += pd[r * groups + gg];
        scale = 1.f / den;                                      // n.pow(-1) (models/attention.py:79)
      } else if (p.gate) scale = __ldg(p.gate + (rowbase + t0 + r) * p.H + h);
      p.out[(rowbase + t0 + r) * p.ldo + (size_t)h * dv + c] = scale * num;
    }
  }
}



// ---- layer forward, column-owner form: thread j of a (b,h) CTA keeps column j of the d x dv state in registers -------------------------------
// out_t[j] = sum_a q_t[a] S_t[a][j] needs no cross-thread reduction, q_t / k_t are broadcast reads of shared memory (one LDS.128 feeds 4 rows of
// every thread), v_t[j] is a per-thread scalar: 2 d FMA per thread-token against d / 2 LDS.128.  The raw q / k / v rows of the NEXT chunk arrive by
// cp.async while the current chunk is computed; phi = elu + 1 is applied in place when the chunk becomes current, and warp 0 then forms the
// normaliser den_t = q_t . sum_{s<=t} k_s of the chunk.  (The thread-per-(row block, column) kernel above spends as many shared-memory reads as FMAs.)
constexpr int LC_TC = 16;

__device__ __forceinline__ void lc_cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d), "l"(gsrc), "r"(n) : "memory");
}

template <int D_>
__global__ void __launch_bounds__(256) linattn_forward_col_kernel(const LinAttnParams p) {
  extern __shared__ __align__(16) float sm[];
  const int dv = p.dv;                                             // == blockDim.x
  const int chunk_floats = LC_TC * (2 * D_ + dv);                  // [q TC x D_ | k TC x D_ | v TC x dv]
  float* buf0 = sm;
  float* buf1 = sm + chunk_floats;
  float* dens = buf1 + chunk_floats;                               // [TC]
  float* ksum_s = dens + LC_TC;                                    // [D_] running sum of phi(k) (normalise)
  const int tid = threadIdx.x, lane = tid & 31, nthr = blockDim.x;
  const int h = blockIdx.x, b = blockIdx.y;
  const size_t rowbase = (size_t)b * p.T;
  const float* qb = p.q + (size_t)h * D_;
  const float* kb = p.k + (size_t)h * D_;
  const float* vb = p.v + (size_t)h * dv;
  float S[D_];
#pragma unroll
  for (int a = 0; a < D_; ++a) S[a] = 0.f;
  for (int a = tid; a < D_; a += nthr) ksum_s[a] = 0.f;

  auto stage = [&](float* dst, int64_t t0) {                       // rows [t0, t0 + TC) -> dst, zero-filled beyond T; 16-byte pieces
    constexpr int QF4 = D_ / 4;
    const int vf4 = dv / 4;
    for (int i = tid; i < LC_TC * QF4; i += nthr) {
      const int r = i / QF4, c = i - r * QF4;
      const bool ok = t0 + r < p.T;
      const size_t row = (rowbase + (ok ? t0 + r : 0)) * p.ld;
      lc_cp_async16(dst + r * D_ + 4 * c, qb + row + 4 * c, ok);
      lc_cp_async16(dst + LC_TC * D_ + r * D_ + 4 * c, kb + row + 4 * c, ok);
    }
    for (int i = tid; i < LC_TC * vf4; i += nthr) {
      const int r = i / vf4, c = i - r * vf4;
      const bool ok = t0 + r < p.T;
      lc_cp_async16(dst + 2 * LC_TC * D_ + r * dv + 4 * c, vb + (rowbase + (ok ? t0 + r : 0)) * p.ld + 4 * c, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  stage(buf0, 0);
  int cur = 0;
  for (int64_t t0 = 0; t0 < p.T; t0 += LC_TC, cur ^= 1) {
    const int tc = (int)min((int64_t)LC_TC, p.T - t0);
    float* cb = cur ? buf1 : buf0;
    float* qs = cb; float* ks = cb + LC_TC * D_; float* vs = cb + 2 * LC_TC * D_;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                               // chunk `cur` landed; every thread is done with the other buffer
    if (t0 + LC_TC < p.T) stage(cur ? buf0 : buf1, t0 + LC_TC);    // next chunk in flight during this one
    if (p.phi_elu) {
      // q rows live at [0, TC D_), k rows at [TC D_, 2 TC D_) whatever tc is: transform both regions in full (on a partial last chunk the rows
      // beyond tc are zero-filled by stage() and never read, so phi of them is harmless; `tc * 2 * D_` here left k rows raw when T % 16 != 0)
      for (int i = tid; i < LC_TC * 2 * D_; i += nthr) cb[i] = elu_f(cb[i]) + 1.f;
      __syncthreads();
    }
    if (p.normalise && tid < 32) {                                 // warp 0: den_t = q_t . ksum_t, ksum_t = ksum_{t-1} + k_t (sequential over the chunk)
      constexpr int PER = (D_ + 31) / 32;
      float kr[PER];
#pragma unroll
      for (int i = 0; i < PER; ++i) kr[i] = (lane + 32 * i < D_) ? ksum_s[lane + 32 * i] : 0.f;
      for (int tt = 0; tt < tc; ++tt) {
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
          const int c = lane + 32 * i;
          if (c < D_) { kr[i] += ks[tt * D_ + c]; part = fmaf(qs[tt * D_ + c], kr[i], part); }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) dens[tt] = part;
      }
#pragma unroll
      for (int i = 0; i < PER; ++i) if (lane + 32 * i < D_) ksum_s[lane + 32 * i] = kr[i];
    }
    if (p.normalise) __syncthreads();
    for (int tt = 0; tt < tc; ++tt) {
      const float vj = vs[tt * dv + tid] * p.kscale;               // (k kscale) v == k (kscale v)
      const float4* k4 = reinterpret_cast<const float4*>(ks + tt * D_);
      const float4* q4 = reinterpret_cast<const float4*>(qs + tt * D_);
      float n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;
#pragma unroll
      for (int a = 0; a < D_ / 4; ++a) {
        const float4 kk = k4[a], qq = q4[a];
        S[4 * a + 0] = fmaf(kk.x, vj, S[4 * a + 0]); n0 = fmaf(qq.x, S[4 * a + 0], n0);
        S[4 * a + 1] = fmaf(kk.y, vj, S[4 * a + 1]); n1 = fmaf(qq.y, S[4 * a + 1], n1);
        S[4 * a + 2] = fmaf(kk.z, vj, S[4 * a + 2]); n2 = fmaf(qq.z, S[4 * a + 2], n2);
        S[4 * a + 3] = fmaf(kk.w, vj, S[4 * a + 3]); n3 = fmaf(qq.w, S[4 * a + 3], n3);
      }
      const float num = (n0 + n1) + (n2 + n3);
      float scale = 1.f;
      if (p.normalise) scale = 1.f / dens[tt];                     // n.pow(-1) (models/attention.py:79)
      else if (p.gate) scale = __ldg(p.gate + (rowbase + t0 + tt) * p.H + h);
      p.out[(rowbase + t0 + tt) * p.ldo + (size_t)h * dv + tid] = scale * num;
    }
  }
}

// ---- softmax attention: normaliser of get_eig_att_softmax and the layer forward, both streaming over key tiles -------------------------------
// Reference: analysis/eval_eig.py:43-95 (the (B,T,T,H) score tensor, its multiplicative mask and the row maximum that therefore includes the
// masked zeros) and SelfAttention.forward, models/attention.py:14-35 (k scaled by 1/sqrt(d) BEFORE the product, additive -10000 mask, softmax).
// A CTA owns SM_TQ query rows of one (b,h); 4 threads share a row and take the keys s = j, j+4, ... of every 32-key tile.  Two passes over the
// key tiles: the row maximum first, then the sums with the FINAL maximum -- the reference subtracts it in float32 and exponentiates in float64,
// which an online rescaling would not reproduce.  Scores are recomputed, not stored: O(T) memory.
constexpr int SM_TQ = 32, SM_TK = 32, SM_THREADS = 128;

__device__ __forceinline__ float sm_dot(const float* __restrict__ qr, const float* __restrict__ kr, int d) {
  float acc = 0.f;
  for (int c = 0; c < d; ++c) acc = fmaf(qr[c], kr[c], acc);
  return acc;
}

__global__ void __launch_bounds__(SM_THREADS) softmax_nu_kernel(const float* __restrict__ q, const float* __restrict__ k, int64_t ld,
                                                                int64_t T, int H, int d, double* __restrict__ nu, float* __restrict__ m) {
  extern __shared__ __align__(16) float sm[];
  const int dp = d + 1;                                            // +1: rows of a tile fall in different banks
  float* qs = sm;                                                  // [SM_TQ][dp]
  float* ks = qs + SM_TQ * dp;                                     // [SM_TK][dp]
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, r = tid >> 2, j = tid & 3;
  const int64_t t = (int64_t)qt * SM_TQ + r;
  const float* qb = q + ((size_t)b * T) * ld + (size_t)h * d;
  const float* kb = k + ((size_t)b * T) * ld + (size_t)h * d;
  for (int i = tid; i < SM_TQ * d; i += SM_THREADS) {
    const int rr = i / d, c = i - rr * d;
    const int64_t tt = (int64_t)qt * SM_TQ + rr;
    qs[rr * dp + c] = tt < T ? __ldg(qb + tt * ld + c) : 0.f;
  }
  float mx = -INFINITY;
  double sum = 0.0;
  for (int pass = 0; pass < 2; ++pass) {
    for (int kt = 0; kt <= qt; ++kt) {
      __syncthreads();
      for (int i = tid; i < SM_TK * d; i += SM_THREADS) {
        const int rr = i / d, c = i - rr * d;
        const int64_t ss = (int64_t)kt * SM_TK + rr;
        ks[rr * dp + c] = ss < T ? __ldg(kb + ss * ld + c) : 0.f;
      }
      __syncthreads();
      if (t < T) {
        for (int s = j; s < SM_TK; s += 4) {
          const int64_t sg = (int64_t)kt * SM_TK + s;
          if (sg <= t) {
            const float dot = sm_dot(qs + r * dp, ks + s * dp, d);
            if (pass == 0) mx = fmaxf(mx, dot);
            else sum += exp((double)(dot - mx));                   // float32 difference, float64 exponential (eval_eig.py:72-77)
          }
        }
      }
    }
    if (pass == 0) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      if (t < T - 1) mx = fmaxf(mx, 0.f);                          // the masked (zeroed) scores take part in torch.max (eval_eig.py:58-62)
    }
  }
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  if (t < T && j == 0) {
    const size_t o = ((size_t)b * T + t) * H + h;
    nu[o] = sum + (double)(T - 1 - t);                             // every masked position contributes exp(0 - 0) = 1
    m[o] = mx;
  }
}

// layer forward: out[t,:] = sum_{s<=t} softmax_s(q_t . (k_s * scale)) v_s.  Pass 0: row maximum and denominator (online), pass 1: P V.
template <int DV>
__global__ void __launch_bounds__(SM_THREADS) softmax_attn_forward_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                                                          int64_t ld, float scale, float* __restrict__ out, int64_t ldo,
                                                                          int64_t T, int H, int d) {
  extern __shared__ __align__(16) float sm[];
  const int dp = d + 1;
  float* qs = sm;                                                  // [SM_TQ][dp]
  float* ks = qs + SM_TQ * dp;                                     // [SM_TK][dp]   k * scale
  float* vs = ks + SM_TK * dp;                                     // [SM_TK][DV]
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, r = tid >> 2, j = tid & 3;
  const int64_t t = (int64_t)qt * SM_TQ + r;
  const float* qb = q + ((size_t)b * T) * ld + (size_t)h * d;
  const float* kb = k + ((size_t)b * T) * ld + (size_t)h * d;
  const float* vb = v + ((size_t)b * T) * ld + (size_t)h * DV;
  for (int i = tid; i < SM_TQ * d; i += SM_THREADS) {
    const int rr = i / d, c = i - rr * d;
    const int64_t tt = (int64_t)qt * SM_TQ + rr;
    qs[rr * dp + c] = tt < T ? __ldg(qb + tt * ld + c) : 0.f;
  }
  float mx = -INFINITY, den = 0.f;
  float acc[DV];
#pragma unroll
  for (int c = 0; c < DV; ++c) acc[c] = 0.f;
  for (int pass = 0; pass < 2; ++pass) {
    for (int kt = 0; kt <= qt; ++kt) {
      __syncthreads();
      for (int i = tid; i < SM_TK * d; i += SM_THREADS) {
        const int rr = i / d, c = i - rr * d;
        const int64_t ss = (int64_t)kt * SM_TK + rr;
        ks[rr * dp + c] = ss < T ? __ldg(kb + ss * ld + c) * scale : 0.f;
      }
      if (pass == 1) {
        for (int i = tid; i < SM_TK * DV; i += SM_THREADS) {
          const int rr = i / DV, c = i - rr * DV;
          const int64_t ss = (int64_t)kt * SM_TK + rr;
          vs[i] = ss < T ? __ldg(vb + ss * ld + c) : 0.f;
        }
      }
      __syncthreads();
      if (t < T) {
        for (int s = j; s < SM_TK; s += 4) {
          const int64_t sg = (int64_t)kt * SM_TK + s;
          if (sg <= t) {
            const float dot = sm_dot(qs + r * dp, ks + s * dp, d);
            if (pass == 0) {
              const float nm = fmaxf(mx, dot);
              den = den * expf(mx - nm) + expf(dot - nm);
              mx = nm;
            } else {
              const float pw = expf(dot - mx) * den;               // den holds 1 / sum in pass 1
#pragma unroll
              for (int c = 0; c < DV; ++c) acc[c] = fmaf(pw, vs[s * DV + c], acc[c]);
            }
          }
        }
      }
    }
    if (pass == 0) {                                               // combine the 4 partial (max, denominator) pairs of the row
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, mx, o), od = __shfl_xor_sync(0xffffffffu, den, o);
        const float nm = fmaxf(mx, om);
        den = (mx == -INFINITY ? 0.f : den * expf(mx - nm)) + (om == -INFINITY ? 0.f : od * expf(om - nm));
        mx = nm;
      }
      den = 1.f / den;
    }
  }
#pragma unroll
  for (int c = 0; c < DV; ++c) {
    acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 1);
    acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 2);
  }
  if (t < T) {
    float* op = out + ((size_t)b * T + t) * ldo + (size_t)h * DV;
#pragma unroll
    for (int c = 0; c < DV; ++c)
      if ((c & 3) == j) op[c] = acc[c];                            // the 4 lanes of the row write interleaved columns
  }
}

}  // namespace eigb200

using namespace eigb200;

extern "C" int eigb200_linattn_nu(void* stream, const float* d_q, const float* d_k, int64_t ld, int64_t B, int64_t T, int H, int d, double* d_nu) {
  EIGB_CHECK_ARG(d_q && d_k && d_nu, "linattn_nu: null pointer");
  EIGB_CHECK_ARG(B > 0 && B <= 65535 && T > 0 && H > 0 && d > 0, "linattn_nu: bad shape");
  EIGB_CHECK_ARG(d <= 32 * NU_DMAX, "linattn_nu: head dim %d > %d", d, 32 * NU_DMAX);
  dim3 grid(H, (unsigned)B);
  linattn_nu_kernel<<<grid, NU_WARPS * 32, 0, (cudaStream_t)stream>>>(d_q, d_k, ld, T, H, d, d_nu);
  EIGB_LAUNCH_CHECK("linattn_nu_kernel");
  return EIGB200_OK;
}

extern "C" int eigb200_linattn_forward(void* stream, const float* d_q, const float* d_k, const float* d_v, int64_t ld,
                                       const float* d_gate, int phi_elu, int normalise, float kscale,
                                       float* d_out, int64_t ldo, int64_t B, int64_t T, int H, int d, int dv) {
  EIGB_CHECK_ARG(d_q && d_k && d_v && d_out, "linattn_forward: null pointer");
  EIGB_CHECK_ARG(B > 0 && B <= 65535 && T > 0 && H > 0 && d > 0 && dv > 0, "linattn_forward: bad shape");
  EIGB_CHECK_ARG(dv <= LA_THREADS && (dv & (dv - 1)) == 0, "linattn_forward: value head dim %d must be a power of two <= 256", dv);
  const int groups = LA_THREADS / dv;
  EIGB_CHECK_ARG(d % groups == 0, "linattn_forward: key head dim %d must be a multiple of %d (= 256 / dv)", d, groups);
  const int ri = d / groups;
  EIGB_CHECK_ARG(ri >= 1 && ri <= 64 && (ri & (ri - 1)) == 0, "linattn_forward: d*dv/256 = %d must be a power of two <= 64", ri);
  LinAttnParams p{d_q, d_k, d_v, ld, d_gate, phi_elu, normalise, kscale, d_out, ldo, T, H, d, dv, nullptr, nullptr, 0, -1, -1, -1};
  {
    // chunked tensor-core form (k9_linattn_mma.cu) for d = dv = 64 (the C1 / C5 heads); EIGB200_LINATTN_FORM=col keeps the recurrent column-owner kernel
    const char* form = getenv("EIGB200_LINATTN_FORM");
    if (!(form && strcmp(form, "col") == 0) && linattn_mma_supported(p)) {
      const cudaError_t e = linattn_mma_launch(p, B, (cudaStream_t)stream);
      if (e != cudaSuccess) return cuda_fail(e, "linattn_mma_kernel");
      return EIGB200_OK;
    }
  }
  const bool al16 = ld % 4 == 0 && (((uintptr_t)d_q | (uintptr_t)d_k | (uintptr_t)d_v) & 15) == 0 && d % 4 == 0;   // cp.async 16-byte pieces
  if ((d == 16 || d == 32 || d == 64 || d == 128) && dv % 32 == 0 && dv <= 256 && al16) {     // column-owner kernel: dv threads per (b,h)
    const size_t smem2 = sizeof(float) * (2 * (size_t)LC_TC * (2 * d + dv) + LC_TC + d);
    dim3 grid2(H, (unsigned)B);
    cudaStream_t st2 = (cudaStream_t)stream;
    // d = 128 with dv >= 128 needs more than the 48 KB a kernel gets without opting in (the launch failed with "invalid argument" before)
#define LC_CASE(D_) case D_: \
      if (smem2 > 48 * 1024) EIGB_CUDA(cudaFuncSetAttribute(linattn_forward_col_kernel<D_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)); \
      linattn_forward_col_kernel<D_><<<grid2, dv, smem2, st2>>>(p); break;
    switch (d) { LC_CASE(16) LC_CASE(32) LC_CASE(64) LC_CASE(128) default: break; }
#undef LC_CASE
    EIGB_LAUNCH_CHECK("linattn_forward_col_kernel");
    return EIGB200_OK;
  }
  const size_t smem = sizeof(float) * ((size_t)LA_TC * (2 * d + dv) + (size_t)LA_TC * groups * dv + (size_t)LA_TC * groups);
  dim3 grid(H, (unsigned)B);
  cudaStream_t st = (cudaStream_t)stream;
#define LA_CASE(RI_) case RI_: \
    if (smem > 48 * 1024) EIGB_CUDA(cudaFuncSetAttribute(linattn_forward_kernel<RI_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    linattn_forward_kernel<RI_><<<grid, LA_THREADS, smem, st>>>(p); break;
  switch (ri) { LA_CASE(1) LA_CASE(2) LA_CASE(4) LA_CASE(8) LA_CASE(16) LA_CASE(32) LA_CASE(64) default: break; }
#undef LA_CASE
  EIGB_LAUNCH_CHECK("linattn_forward_kernel");
  return EIGB200_OK;
}

extern "C" int eigb200_linattn_conv_fusable(const float* d_q, const float* d_k, const float* d_v, int64_t ld, const float* d_out, int64_t ldo, int d, int dv, int kconv) {
  LinAttnParams p{d_q, d_k, d_v, ld, nullptr, 1, 0, 1.f, const_cast<float*>(d_out), ldo, 1, 1, d, dv, d_q, d_q, kconv, 0, 0, 0};
  return linattn_mma_supported(p) ? 1 : 0;
}

extern "C" int eigb200_linattn_forward_conv(void* stream, const float* d_q, const float* d_k, const float* d_v, int64_t ld,
                                            const float* d_gate, int phi_elu, int normalise, float kscale,
                                            const float* d_conv_w, const float* d_conv_b, int kconv, int conv_ch_q, int conv_ch_k, int conv_ch_v,
                                            float* d_out, int64_t ldo, int64_t B, int64_t T, int H, int d, int dv) {
  EIGB_CHECK_ARG(d_q && d_k && d_v && d_out && d_conv_w && d_conv_b, "linattn_forward_conv: null pointer");
  EIGB_CHECK_ARG(B > 0 && B <= 65535 && T > 0 && H > 0 && d > 0 && dv > 0, "linattn_forward_conv: bad shape");
  LinAttnParams p{d_q, d_k, d_v, ld, d_gate, phi_elu, normalise, kscale, d_out, ldo, T, H, d, dv, d_conv_w, d_conv_b, kconv, conv_ch_q, conv_ch_k, conv_ch_v};
  EIGB_CHECK_ARG(linattn_mma_supported(p), "linattn_forward_conv: needs d = dv = 64, conv taps <= 4, 16-byte aligned q / k / v rows (ld %% 4 == 0) and an 8-byte aligned output "
                 "(got d %d dv %d k %d ld %lld): run eigb200_conv_silu + eigb200_linattn_forward instead", d, dv, kconv, (long long)ld);
  const cudaError_t e = linattn_mma_launch(p, B, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "linattn_mma_kernel<conv>");
  return EIGB200_OK;
}

extern "C" int eigb200_softmax_nu(void* stream, const float* d_q, const float* d_k, int64_t ld, int64_t B, int64_t T, int H, int d,
                                  double* d_nu, float* d_m) {
  EIGB_CHECK_ARG(d_q && d_k && d_nu && d_m, "softmax_nu: null pointer");
  EIGB_CHECK_ARG(B > 0 && B <= 65535 && T > 0 && H > 0 && H <= 65535 && d > 0, "softmax_nu: bad shape");
  const size_t smem = sizeof(float) * (size_t)(SM_TQ + SM_TK) * (d + 1);
  EIGB_CHECK_ARG(smem <= 200 * 1024, "softmax_nu: head dim %d too large", d);
  if (smem > 48 * 1024) EIGB_CUDA(cudaFuncSetAttribute(softmax_nu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((T + SM_TQ - 1) / SM_TQ), H, (unsigned)B);
  softmax_nu_kernel<<<grid, SM_THREADS, smem, (cudaStream_t)stream>>>(d_q, d_k, ld, T, H, d, d_nu, d_m);
  EIGB_LAUNCH_CHECK("softmax_nu_kernel");
  return EIGB200_OK;
}

extern "C" int eigb200_softmax_attn_forward(void* stream, const float* d_q, const float* d_k, const float* d_v, int64_t ld, float scale,
                                            float* d_out, int64_t ldo, int64_t B, int64_t T, int H, int d, int dv) {
  EIGB_CHECK_ARG(d_q && d_k && d_v && d_out, "softmax_attn_forward: null pointer");
  EIGB_CHECK_ARG(B > 0 && B <= 65535 && T > 0 && H > 0 && H <= 65535 && d > 0, "softmax_attn_forward: bad shape");
  EIGB_CHECK_ARG(dv == 16 || dv == 32 || dv == 64 || dv == 128, "softmax_attn_forward: value head dim %d must be 16, 32, 64 or 128", dv);
  const size_t smem = sizeof(float) * ((size_t)(SM_TQ + SM_TK) * (d + 1) + (size_t)SM_TK * dv);
  EIGB_CHECK_ARG(smem <= 200 * 1024, "softmax_attn_forward: head dims %d/%d too large", d, dv);
  dim3 grid((unsigned)((T + SM_TQ - 1) / SM_TQ), H, (unsigned)B);
  cudaStream_t st = (cudaStream_t)stream;
#define SMF_CASE(DV_) case DV_: \
    if (smem > 48 * 1024) EIGB_CUDA(cudaFuncSetAttribute(softmax_attn_forward_kernel<DV_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    softmax_attn_forward_kernel<DV_><<<grid, SM_THREADS, smem, st>>>(d_q, d_k, d_v, ld, scale, d_out, ldo, T, H, d); break;
  switch (dv) { SMF_CASE(16) SMF_CASE(32) SMF_CASE(64) SMF_CASE(128) default: break; }
#undef SMF_CASE
  EIGB_LAUNCH_CHECK("softmax_attn_forward_kernel");
  return EIGB200_OK;
}
