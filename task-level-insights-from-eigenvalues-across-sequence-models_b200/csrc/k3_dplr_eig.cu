// k3_dplr_eig.cu -- K3: S4 DPLR discretisation and a batched small-N nonsymmetric eigensolver (one warp per matrix).
//
// Reference operators:
//   discrete_DPLR (A-bar only)            analysis/eval_eig.py:254-274, models/s4.py:16-36
//   np.linalg.eigvals(Ad)                 analysis/eval_eig.py:296
// NumPy's linalg promotes complex64 input to complex128 (zgeev) and casts the eigenvalues back to complex64, so the solver here
// works in float64 on the complex64 matrix it is given: balancing (LAPACK xGEBAL-style power-of-two row/column scaling),
// Householder reduction to upper Hessenberg form, then the explicit single-shift complex QR iteration (Wilkinson shift,
// accumulated shifts as in EISPACK COMQR, exceptional shifts at iterations 10 and 20), eigenvalues only -- rotations are applied
// to the active window [l, en] alone.  One warp owns one matrix in shared memory ((N+1)-padded rows of double2); lanes span
// columns for row operations and rows for column operations; all control flow is warp uniform.
// Bound: latency / FP64 ALU (the QR sweep is a serial chain of 2*(en-l) rotations); no HBM roofline applies.
#include "common.cuh"

namespace eigb200 {

typedef double2 cd;
__device__ __forceinline__ cd cmk(double r, double i) { return make_double2(r, i); }
__device__ __forceinline__ cd cadd(cd a, cd b) { return cmk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd csub(cd a, cd b) { return cmk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cd cmul(cd a, cd b) { return cmk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cd cmulc(cd a, cd b) { return cmk(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }   // a * conj(b)
__device__ __forceinline__ cd cscale(cd a, double s) { return cmk(a.x * s, a.y * s); }
__device__ __forceinline__ cd cconj(cd a) { return cmk(a.x, -a.y); }
__device__ __forceinline__ double cabs1(cd a) { return fabs(a.x) + fabs(a.y); }
__device__ __forceinline__ double cabs2(cd a) { return a.x * a.x + a.y * a.y; }
__device__ __forceinline__ double cabsd(cd a) { return hypot(a.x, a.y); }
__device__ __forceinline__ cd cdiv(cd a, cd b) {
  // Smith's algorithm
  if (fabs(b.x) >= fabs(b.y)) { const double r = b.y / b.x, d = b.x + b.y * r; return cmk((a.x + a.y * r) / d, (a.y - a.x * r) / d); }
  const double r = b.x / b.y, d = b.x * r + b.y; return cmk((a.x * r + a.y) / d, (a.y * r - a.x) / d);
}
__device__ __forceinline__ cd csqrt_d(cd z) {
  const double m = cabsd(z);
  if (m == 0.0) return cmk(0.0, 0.0);
  double re = sqrt(0.5 * (m + fabs(z.x)));
  double im = 0.5 * z.y / re;
  if (z.x < 0.0) { const double t = re; re = fabs(im); im = copysign(t, z.y); }
  return cmk(re, im);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int EIG_MAXN = 64;
constexpr int EIG_MAXITS = 30;

// One warp, one matrix.  H: n x ld (ld = n + 1) double2 in shared memory; vec: n double2; rc: n double.
__device__ void warp_eigvals(cd* H, cd* vec, double* rc, int n, int ld, int lane, float2* eig_out, int* info_out) {
#define HH(i, j) H[(i) * ld + (j)]
  // ---- 1. balancing (sequential over i like xGEBAL; power-of-two factors are exact) ------------------------------------
  for (int sweep = 0; sweep < 8; ++sweep) {
    bool changed = false;
    for (int i = 0; i < n; ++i) {
      double c = 0.0, r = 0.0;
      for (int j = lane; j < n; j += 32) if (j != i) { c += cabs1(HH(j, i)); r += cabs1(HH(i, j)); }
      c = warp_sum(c); r = warp_sum(r);
      if (c == 0.0 || r == 0.0) continue;
      double f = 1.0, ci = c;
      const double s = c + r;
      double g = r * 0.5;
      while (ci < g) { f *= 2.0; ci *= 4.0; }
      g = r * 2.0;
      while (ci >= g) { f *= 0.5; ci *= 0.25; }
      if ((ci + r) / f < 0.95 * s) {
        const double fi = 1.0 / f;
        for (int j = lane; j < n; j += 32) if (j != i) { HH(i, j) = cscale(HH(i, j), fi); HH(j, i) = cscale(HH(j, i), f); }
        changed = true;
      }
      __syncwarp();
    }
    if (!changed) break;
  }
  // ---- 2. Householder reduction to upper Hessenberg form ------------------------------------------------------------------
  for (int k = 0; k + 2 < n; ++k) {
    const int m = n - k - 1;
    double nrm2 = 0.0;
    for (int i = lane; i < m; i += 32) nrm2 += cabs2(HH(k + 1 + i, k));
    nrm2 = warp_sum(nrm2);
    const cd x0 = HH(k + 1, k);
    const double ax0sq = cabs2(x0);
    if (nrm2 - ax0sq <= 0.0) continue;                      // nothing below the subdiagonal
    const double nrm = sqrt(nrm2), ax0 = sqrt(ax0sq);
    const cd phase = ax0 > 0.0 ? cscale(x0, 1.0 / ax0) : cmk(1.0, 0.0);
    const cd beta = cscale(phase, -nrm);
    const cd v0 = csub(x0, beta);
    const double tau = 2.0 / (nrm2 - ax0sq + cabs2(v0));
    for (int i = lane; i < m; i += 32) vec[i] = i == 0 ? v0 : HH(k + 1 + i, k);
    __syncwarp();
    // left: H[k+1:, k:] -= tau * v (v^H H[k+1:, k:])
    for (int j = k + lane; j < n; j += 32) {
      cd w = cmk(0.0, 0.0);
      for (int i = 0; i < m; ++i) w = cadd(w, cmulc(HH(k + 1 + i, j), vec[i]));          // conj(v_i) * H
      w = cscale(w, tau);
      for (int i = 0; i < m; ++i) HH(k + 1 + i, j) = csub(HH(k + 1 + i, j), cmul(vec[i], w));
    }
    __syncwarp();
    // right: H[:, k+1:] -= tau * (H[:, k+1:] v) v^H
    for (int r = lane; r < n; r += 32) {
      cd w = cmk(0.0, 0.0);
      for (int i = 0; i < m; ++i) w = cadd(w, cmul(HH(r, k + 1 + i), vec[i]));
      w = cscale(w, tau);
      for (int i = 0; i < m; ++i) HH(r, k + 1 + i) = csub(HH(r, k + 1 + i), cmulc(w, vec[i]));
    }
    __syncwarp();
    for (int i = lane; i < m; i += 32) HH(k + 1 + i, k) = i == 0 ? beta : cmk(0.0, 0.0);
    __syncwarp();
  }
  // ---- 3. single-shift QR on the Hessenberg matrix -------------------------------------------------------------------------
  const double eps = 2.220446049250313e-16;
  double anorm = 0.0;
  for (int idx = lane; idx < n * n; idx += 32) { const int i = idx / n, j = idx - i * n; if (j + 1 >= i) anorm += cabs1(HH(i, j)); }
  anorm = warp_sum(anorm);
  cd t = cmk(0.0, 0.0);
  int en = n - 1, its = 0, info = 0;
  while (en >= 0) {
    // largest l <= en with a negligible subdiagonal H[l][l-1] (or l = 0)
    int l = 0;
    for (int base = en; base > 0; base -= 32) {
      const int cand = base - lane;
      bool small = false;
      if (cand > 0) {
        double s = cabs1(HH(cand - 1, cand - 1)) + cabs1(HH(cand, cand));
        if (s == 0.0) s = anorm;
        small = cabs1(HH(cand, cand - 1)) <= eps * s;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, small);
      if (bal) { l = base - (__ffs(bal) - 1); break; }
    }
    if (l == en) {
      if (lane == 0) { const cd e = cadd(HH(en, en), t); eig_out[en] = make_float2((float)e.x, (float)e.y); }
      --en; its = 0;
      continue;
    }
    if (its >= EIG_MAXITS) { info = en + 1; break; }
    cd sh;
    if (its == 10 || its == 20) {
      sh = cmk(fabs(HH(en, en - 1).x) + (en >= 2 ? fabs(HH(en - 1, en - 2).x) : 0.0), 0.0);
    } else {
      const cd a = HH(en - 1, en - 1), b = HH(en - 1, en), c = HH(en, en - 1), d = HH(en, en);
      sh = d;
      const cd bc = cmul(b, c);
      if (bc.x != 0.0 || bc.y != 0.0) {
        const cd hf = cscale(csub(a, d), 0.5);
        cd disc = csqrt_d(cadd(cmul(hf, hf), bc));
        if (hf.x * disc.x + hf.y * disc.y < 0.0) disc = cmk(-disc.x, -disc.y);
        sh = csub(d, cdiv(bc, cadd(hf, disc)));
      }
    }
    ++its;
    __syncwarp();
    for (int i = lane; i <= en; i += 32) HH(i, i) = csub(HH(i, i), sh);
    t = cadd(t, sh);
    __syncwarp();
    // H <- G_{en-1} .. G_l H G_l^H .. G_{en-1}^H.  EISPACK runs a row pass (H <- G_k H, columns k .. en) and then a column pass (H <- H G_k^H, rows
    // l .. min(k+1, en)); here step k does the row rotation G_k AND the column rotation G_{k-2}^H: the two touch disjoint elements (columns >= k against
    // columns k-2, k-1), rows <= k-1 have seen all their row rotations by then, and every element still receives its updates in EISPACK's order, so the
    // results are bit-identical -- but the serial chain is half as long, and the column update's loads run under the hypot / divisions of the next rotation.
    for (int k = l; k < en + 2; ++k) {
      const bool do_row = k < en;
      double cth = 1.0; cd sn = cmk(0.0, 0.0);
      if (do_row) {
        // c = |a| / r, s = (a / |a|) conj(b) / r, r = sqrt(|a|^2 + |b|^2): the step's serial chain.  Two independent rsqrt (|a|^2 and r^2) and products
        // instead of three hypot and three divisions (~500 -> ~100 cycles per step); the squares cannot leave the double range for a balanced matrix whose
        // negligible subdiagonals have been deflated (|.| >= eps * local scale), and (c, s) are unitary to rounding as before.
        const cd a = HH(k, k), b = HH(k + 1, k);
        const double aa2 = cabs2(a), ab2 = cabs2(b);
        const double r2 = aa2 + ab2;
        if (r2 != 0.0) {
          if (aa2 == 0.0) { cth = 0.0; sn = cscale(cconj(b), rsqrt(ab2)); }
          else {
            const double ia = rsqrt(aa2), ir = rsqrt(r2);
            cth = aa2 * ia * ir;                                                           // |a| / r
            sn = cscale(cmulc(a, b), ia * ir);                                             // a conj(b) / (|a| r)
          }
        }
      }
      __syncwarp();                                          // everyone has read a, b before they are overwritten
      if (do_row) {
        for (int j = k + lane; j <= en; j += 32) {
          const cd top = HH(k, j), bot = HH(k + 1, j);
          HH(k, j) = cadd(cscale(top, cth), cmul(sn, bot));
          HH(k + 1, j) = j == k ? cmk(0.0, 0.0) : csub(cscale(bot, cth), cmul(cconj(sn), top));
        }
        if (lane == 0) { rc[k] = cth; vec[k] = sn; }
      }
      const int kc = k - 2;
      if (kc >= l) {                                         // column rotation kc: its (c, s) were stored two steps ago
        const double c2 = rc[kc]; const cd s2 = vec[kc];
        const int hi = min(kc + 1, en);
        for (int i = l + lane; i <= hi; i += 32) {
          const cd c0 = HH(i, kc), c1 = HH(i, kc + 1);
          HH(i, kc) = cadd(cscale(c0, c2), cmul(cconj(s2), c1));
          HH(i, kc + 1) = csub(cscale(c1, c2), cmul(s2, c0));
        }
      }
      __syncwarp();
    }
  }
  if (info != 0) {                                           // unconverged: report what is on the diagonal
    for (int i = lane; i <= en; i += 32) { const cd e = cadd(HH(i, i), t); eig_out[i] = make_float2((float)e.x, (float)e.y); }
  }
  if (lane == 0) *info_out = info;
#undef HH
}

template <int WPB>
__global__ void __launch_bounds__(WPB * 32) eigvals_kernel(const float2* __restrict__ A, int64_t nmat, int n, float2* __restrict__ eig, int* __restrict__ info) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = n + 1;
  const size_t per_warp = (((size_t)n * ld + n) * sizeof(cd) + (size_t)n * sizeof(double) + 15) & ~(size_t)15;
  unsigned char* base = sm_raw + warp * per_warp;
  cd* H = reinterpret_cast<cd*>(base);
  cd* vec = H + (size_t)n * ld;
  double* rc = reinterpret_cast<double*>(vec + n);
  for (int64_t mat = (int64_t)blockIdx.x * WPB + warp; mat < nmat; mat += (int64_t)gridDim.x * WPB) {
    const float2* Am = A + mat * n * n;
    for (int idx = lane; idx < n * n; idx += 32) { const int i = idx / n, j = idx - i * n; const float2 v = Am[idx]; H[i * ld + j] = cmk((double)v.x, (double)v.y); }
    __syncwarp();
    warp_eigvals(H, vec, rc, n, ld, lane, eig + mat * n, info + mat);
    __syncwarp();
  }
}

// A-bar = A1 A0 in closed form (rank-one structure):  Abar_ij = delta_ij d_i a_i - c (d_i P_i) conj(Q_j) (1 + d_j a_j),
//   a = 2/step + Lambda, d = 1/(2/step - Lambda), c = 1/(1 + sum_i conj(Q_i) d_i P_i)      (float64 internally, complex64 out)
__global__ void __launch_bounds__(256) dplr_abar_kernel(const float2* __restrict__ Lam, const float2* __restrict__ Pv, const float2* __restrict__ Qv,
                                                        const float* __restrict__ step, int n, float2* __restrict__ Abar) {
  __shared__ cd dP[EIG_MAXN], qf[EIG_MAXN], da[EIG_MAXN];
  __shared__ cd kappa_s;
  const int64_t mat = blockIdx.x;
  const double two = 2.0 / (double)step[mat];
  const float2* L = Lam + mat * n; const float2* P = Pv + mat * n; const float2* Q = Qv + mat * n;
  if (threadIdx.x < n) {
    const int i = threadIdx.x;
    const cd lam = cmk(L[i].x, L[i].y), p = cmk(P[i].x, P[i].y), q = cmk(Q[i].x, Q[i].y);
    const cd d = cdiv(cmk(1.0, 0.0), cmk(two - lam.x, -lam.y));
    const cd a = cmk(two + lam.x, lam.y);
    dP[i] = cmul(d, p);
    da[i] = cmul(d, a);
    qf[i] = cmul(cconj(q), cadd(cmk(1.0, 0.0), da[i]));       // conj(Q_j) (1 + d_j a_j)
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    cd k = cmk(0.0, 0.0);
    for (int i = 0; i < n; ++i) k = cadd(k, cmulc(dP[i], cmk(Q[i].x, Q[i].y)));   // conj(Q_i) d_i P_i
    kappa_s = cdiv(cmk(1.0, 0.0), cadd(cmk(1.0, 0.0), k));
  }
  __syncthreads();
  const cd c = kappa_s;
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int i = idx / n, j = idx - i * n;
    cd v = cmul(cmul(c, dP[i]), qf[j]);
    v = cmk(-v.x, -v.y);
    if (i == j) v = cadd(v, da[i]);
    Abar[mat * n * n + idx] = make_float2((float)v.x, (float)v.y);
  }
}

}  // namespace eigb200

using namespace eigb200;

extern "C" int eigb200_dplr_abar(void* stream, const float* d_Lambda, const float* d_P, const float* d_Q, const float* d_step,
                                 int64_t nmat, int N, float* d_Abar) {
  EIGB_CHECK_ARG(d_Lambda && d_P && d_Q && d_step && d_Abar, "dplr_abar: null pointer");
  EIGB_CHECK_ARG(nmat > 0 && N >= 1 && N <= EIG_MAXN, "dplr_abar: need 1 <= N <= %d, got %d", EIG_MAXN, N);
  EIGB_CHECK_ARG(nmat < (1LL << 31), "dplr_abar: too many matrices");
  dplr_abar_kernel<<<(unsigned)nmat, 256, 0, (cudaStream_t)stream>>>((const float2*)d_Lambda, (const float2*)d_P, (const float2*)d_Q, d_step, N, (float2*)d_Abar);
  EIGB_LAUNCH_CHECK("dplr_abar_kernel");
  return EIGB200_OK;
}

extern "C" int eigb200_eigvals_c64(void* stream, float* d_A, int64_t nmat, int N, float* d_eig, int32_t* d_info) {
  EIGB_CHECK_ARG(d_A && d_eig && d_info, "eigvals_c64: null pointer");
  EIGB_CHECK_ARG(nmat > 0 && N >= 1 && N <= EIG_MAXN, "eigvals_c64: need 1 <= N <= %d, got %d", EIG_MAXN, N);
  const size_t per_warp = (((size_t)N * (N + 1) + N) * sizeof(double2) + (size_t)N * sizeof(double) + 15) & ~(size_t)15;
  cudaStream_t st = (cudaStream_t)stream;
  // warps per CTA: as many matrices as fit in ~200 KB of shared memory, at most 8
  int wpb = (int)((224 * 1024) / per_warp);
  if (wpb > 8) wpb = 8;
  if (wpb >= 4 && wpb < 8) wpb = 4;
  if (wpb < 1) wpb = 1;                                              // N = 64: 3 matrices (3 x 66.5 KB) per SM -- the solver is latency-bound, every resident warp counts
  const size_t smem = per_warp * wpb;
  int64_t grid = (nmat + wpb - 1) / wpb;
  const int64_t cap = (int64_t)num_sms() * 4;
  if (grid > cap) grid = cap;
#define EIG_LAUNCH(W_)                                                                                                   \
  do {                                                                                                                   \
    if (smem > 48 * 1024) EIGB_CUDA(cudaFuncSetAttribute(eigvals_kernel<W_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    eigvals_kernel<W_><<<(unsigned)grid, W_ * 32, smem, st>>>((const float2*)d_A, nmat, N, (float2*)d_eig, d_info);       \
  } while (0)
  switch (wpb) { case 8: EIG_LAUNCH(8); break; case 4: EIG_LAUNCH(4); break; case 3: EIG_LAUNCH(3); break; case 2: EIG_LAUNCH(2); break; default: EIG_LAUNCH(1); break; }
#undef EIG_LAUNCH
  EIGB_LAUNCH_CHECK("eigvals_kernel");
  return EIGB200_OK;
}
