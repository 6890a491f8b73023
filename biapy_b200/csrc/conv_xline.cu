// x-line convolution: 3x3x3 Conv3d of the 16-output-channel layers at W = 128, with the preceding GroupNorm-apply + SiLU fused
// into the operand path (reference order blocks.py:1304-1378: GN(in) -> act -> conv; blocks.py:148-160 for the statistics).
//
// Why another kernel family.  The x-folded kernels (conv_umma.cu) feed tcgen05.mma from shared memory: at
// Cout = 16 every MMA re-reads its 4 KB A tile for 32 cycles of math, half of which multiplies block-Toeplitz zeros, and each
// input element lands five times (x window, z halo, dy stages) -- which is also what made a GN + SiLU transform on the operand
// path unaffordable.  Here one GEMM row is ONE voxel and the A operand lives in TENSOR MEMORY:
//
//   D[x][(s, co)] += A_dx[x][ci] * B_r[dy][dx][(s, co)][ci]        M = 128 voxels of one input line (y_in, z_in), K = Cin,
//                                                                   N = 48 = three output z-planes x 16 output channels
//
//   * a line of 128 voxels x Cin channels is contiguous in a dense channels-last tensor: it arrives with one 1-D bulk copy
//     (cp.async.bulk, no tensor map, two lines per request) in a raw shared-memory ring;
//   * four transform warps (thread = voxel = TMEM lane) read their voxel once, apply y = silu(x * scale[n,c] + shift[n,c])
//     (the GroupNorm-apply + activation pass of the unfused path, same arithmetic and the same rounding to the engine dtype),
//     optionally write the activated tensor out for the backward pass, and store the voxel into tensor memory three times:
//     as is and shifted by one lane up / down (warp shuffles + a 32-byte exchange at warp borders, zeros at the line ends =
//     'same' padding in x).  An input element is transformed once per CTA that needs it, not five times;
//   * tcgen05.mma with A in TMEM costs N/2 cycles (no 4 KB shared-memory fetch per MMA): the B operand is 1.5 KB per MMA;
//   * no structural zeros: dx is a choice of A copy, dy a choice of output line (input line i feeds output lines i-2, i-1, i
//     of the band), dz is the N dimension: block s of the 48 columns belongs to output plane z_out = s (mod 3), and the
//     weights are packed in three rotations r = z_in mod 3 so that block s always receives tap dz = (r + 1 - s) mod 3.
//     Accumulators (BY lines x 48 columns) stay in TMEM while the CTA walks along z; after input plane p the block of plane
//     p - 1 is complete: the epilogue reads 16 columns, zeroes them (every MMA accumulates), adds the bias (+ the old value
//     for the residual / shared-gradient form), rounds, stores 32 bytes per thread = 1 KB per warp, and keeps the channel
//     sums of the stored values for the following normalisation (conv -> norm, == b200_channel_sums of the output).
//
// Work unit = (sample, z chunk, band of BY output lines); planes z0-1 .. zhi and lines y0-1 .. y0+BY are read (halo).
// Warps: 0-3 epilogue, 4-7 transform, 8 MMA issue (one elected lane), 9 bulk-copy issue (one elected lane).
#include "umma.cuh"

namespace b200 {
namespace sm100 {

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"((uint64_t)src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  const uint32_t z = 0;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(z)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] += A[tmem] * B[smem]: A = 128 lanes x 8 columns (16 K-elements of 16 bits, two per column)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ float xl_tanh(float x) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// FUSE 1: exact chain of b200_scale_shift_act (act_fwd, B200_ACT_SILU); FUSE 2: one-MUFU chain of b200_scale_shift_silu_fast
template <int FUSE>
__device__ __forceinline__ float xl_silu(float z) {
  if (FUSE == 2) return z * fmaf(0.5f, xl_tanh(0.5f * z), 0.5f);
  return __fdividef(z, 1.f + __expf(-z));
}

struct XlineParams {
  int n, d, h;
  int bands, zchunks, zc, units;
  long long xsh_b, xsd_b, xsn_b;   // byte strides of the input: line, plane, sample (voxels of a line are dense)
  long long ash_b, asd_b, asn_b;   // same for the optional activated copy
  long long ysw, ysh, ysd, ysn;    // element strides of the output
  int accumulate;
  uint32_t idesc;
};

constexpr uint32_t kXlTileBytes = 48u * 32u;   // one B tile: 48 rows (s, co) x 16 ci, SWIZZLE_32B K-major

__device__ __forceinline__ int xl_mod3(int v) { return ((v % 3) + 3) % 3; }

template <typename T, int KS, int BY, int FUSE>
__global__ void __launch_bounds__(320, 1)
conv_fprop_xline_kernel(const T* __restrict__ x, const T* __restrict__ wpk, const float* __restrict__ bias, T* __restrict__ y,
                        T* __restrict__ a_out, const float* __restrict__ scale, const float* __restrict__ shift,
                        double* __restrict__ stats, const XlineParams p) {
  constexpr int NR = KS == 1 ? 8 : 3;            // raw ring: slots of two lines
  constexpr int NA = KS == 1 ? 5 : 4;            // operand ring in tensor memory: slots of one line (three shifted copies)
  constexpr int PAIRS = (BY + 2) / 2;
  constexpr uint32_t LINE = 4096u * KS;
  constexpr uint32_t ACOLS = 24u * KS;
  constexpr uint32_t ACC = (uint32_t)BY * 48u;
  constexpr uint32_t BBYTES = 27u * KS * kXlTileBytes;
  constexpr int W = 8 * KS;                      // 32-bit words per voxel
  static_assert(BY % 2 == 0 && ACC + NA * ACOLS <= 512, "tensor memory budget");

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_bar[2 * NR + 2 * NA + 2 * BY + 1];
  __shared__ uint32_t s_tmem;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sm_b = smem0, sm_raw = smem0 + BBYTES, sm_ex = sm_raw + NR * 2u * LINE;
  const uint32_t bar0 = smem_u32(s_bar);
  const uint32_t raw_full = bar0, raw_free = raw_full + 8 * NR, a_full = raw_free + 8 * NR, a_free = a_full + 8 * NA,
                 acc_full = a_free + 8 * NA, acc_free = acc_full + 8 * BY, w_full = acc_free + 8 * BY;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NR; ++i) { mbar_init(raw_full + 8 * i, 1); mbar_init(raw_free + 8 * i, 4); }
    for (int i = 0; i < NA; ++i) { mbar_init(a_full + 8 * i, 4); mbar_init(a_free + 8 * i, 1); }
    for (int i = 0; i < BY; ++i) { mbar_init(acc_full + 8 * i, 1); mbar_init(acc_free + 8 * i, 4); }
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(&s_tmem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  auto decode = [&](int u, int& n, int& z0, int& zhi, int& y0) {
    const int band = u % p.bands;
    int t = u / p.bands;
    const int zk = t % p.zchunks;
    n = t / p.zchunks;
    y0 = band * BY;
    z0 = zk * p.zc;
    zhi = z0 + p.zc < p.d ? z0 + p.zc : p.d;
  };

  if (warp == 9) {
    // ===================================================================== bulk-copy issue
    if (elect_one()) {
      mbar_expect_tx(w_full, BBYTES);
      bulk_g2s(sm_b, wpk, BBYTES, w_full);
      int rs = 0;
      uint32_t rph = 0;
      const char* xb = reinterpret_cast<const char*>(x);
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        int n, z0, zhi, y0;
        decode(u, n, z0, zhi, y0);
        for (int pz = z0 - 1; pz <= zhi; ++pz) {
          if ((unsigned)pz >= (unsigned)p.d) continue;
          const char* plane = xb + (long long)n * p.xsn_b + (long long)pz * p.xsd_b;
#pragma unroll 1
          for (int pr = 0; pr < PAIRS; ++pr) {
            int lo = 2 * pr, hi = 2 * pr + 2;
            if ((unsigned)(y0 - 1 + lo) >= (unsigned)p.h) ++lo;
            if ((unsigned)(y0 - 1 + hi - 1) >= (unsigned)p.h) --hi;
            if (lo >= hi) continue;
            mbar_wait(raw_free + 8 * rs, rph ^ 1);
            const uint32_t bytes = (uint32_t)(hi - lo) * LINE;
            mbar_expect_tx(raw_full + 8 * rs, bytes);
            bulk_g2s(sm_raw + (uint32_t)rs * 2u * LINE + (uint32_t)(lo - 2 * pr) * LINE,
                     plane + (long long)(y0 - 1 + lo) * p.xsh_b, bytes, raw_full + 8 * rs);
            if (++rs == NR) { rs = 0; rph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 8) {
    // ===================================================================== MMA issue
    if (elect_one()) {
      const uint32_t idesc = in_reg(p.idesc);
      const uint32_t b_hi = (256u >> 4) | (1u << 14) | ((uint32_t)kSwizzle32 << 29);
      const uint32_t b_lo0 = ((sm_b >> 4) & 0x3FFFu) | (1u << 16);
      mbar_wait(w_full, 0);
      int slot = 0;
      uint32_t aph = 0, pc = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        int n, z0, zhi, y0;
        decode(u, n, z0, zhi, y0);
        for (int pz = z0 - 1; pz <= zhi; ++pz, ++pc) {
          const bool pv = (unsigned)pz < (unsigned)p.d;
          const uint32_t b_r = b_lo0 + (uint32_t)xl_mod3(pz) * (9u * KS * (kXlTileBytes >> 4));
#pragma unroll
          for (int i = 0; i < BY + 2; ++i) {
            if (i < BY) {
              mbar_wait(acc_free + 8 * i, pc & 1);
              tc_fence_after();
            }
            const bool lv = pv && (unsigned)(y0 - 1 + i) < (unsigned)p.h;
            if (lv) {
              mbar_wait(a_full + 8 * slot, aph);
              tc_fence_after();
              const uint32_t a_base = tmem + ACC + (uint32_t)slot * ACOLS;
#pragma unroll
              for (int oo = i - 2; oo <= i; ++oo) {
                if (oo < 0 || oo >= BY) continue;
                const int dy = i - oo;
                const uint32_t d_t = tmem + (uint32_t)oo * 48u;
#pragma unroll
                for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                  for (int k = 0; k < KS; ++k)
                    umma_f16_ts(d_t, a_base + (uint32_t)(dx * KS + k) * 8u,
                                b_r + (uint32_t)((dy * 3 + dx) * KS + k) * (kXlTileBytes >> 4), b_hi, idesc, 1u);
                if (oo == i - 2) umma_commit(acc_full + 8 * oo);
              }
              umma_commit(a_free + 8 * slot);
              if (++slot == NA) { slot = 0; aph ^= 1; }
            } else if (i >= 2) {
              umma_commit(acc_full + 8 * (i - 2));
            }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== transform: raw line -> three operand copies in TMEM
    const int q = warp & 3;
    const int xv = q * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16) + ACC;
    int rs = 0, slot = 0;
    uint32_t rph = 0, aph = 0, par = 0;
    float sc[KS == 1 ? 16 : 1], sh[KS == 1 ? 16 : 1];
    const char* ab = reinterpret_cast<const char*>(a_out);
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      int n, z0, zhi, y0;
      decode(u, n, z0, zhi, y0);
      if constexpr (FUSE != 0 && KS == 1) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          sc[c] = __ldg(scale + (long long)n * 16 + c);
          sh[c] = __ldg(shift + (long long)n * 16 + c);
        }
      }
      for (int pz = z0 - 1; pz <= zhi; ++pz) {
        if ((unsigned)pz >= (unsigned)p.d) continue;
#pragma unroll 1
        for (int pr = 0; pr < PAIRS; ++pr) {
          int lo = 2 * pr, hi = 2 * pr + 2;
          if ((unsigned)(y0 - 1 + lo) >= (unsigned)p.h) ++lo;
          if ((unsigned)(y0 - 1 + hi - 1) >= (unsigned)p.h) --hi;
          if (lo >= hi) continue;
          mbar_wait(raw_full + 8 * rs, rph);
          for (int i = lo; i < hi; ++i) {
            uint32_t v[W], lf[W], rt[W];
            const uint32_t src = sm_raw + (uint32_t)rs * 2u * LINE + (uint32_t)(i - 2 * pr) * LINE + (uint32_t)xv * (32u * KS);
#pragma unroll
            for (int j = 0; j < W / 4; ++j)
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(v[4 * j]), "=r"(v[4 * j + 1]), "=r"(v[4 * j + 2]), "=r"(v[4 * j + 3])
                           : "r"(src + 16u * j));
            if (FUSE) {
#pragma unroll
              for (int j = 0; j < W; ++j) {
                Pack<T, 2> e = *reinterpret_cast<Pack<T, 2>*>(&v[j]);
                float s0, s1, h0, h1;
                if constexpr (KS == 1) { s0 = sc[2 * j]; s1 = sc[2 * j + 1]; h0 = sh[2 * j]; h1 = sh[2 * j + 1]; }
                else {
                  s0 = __ldg(scale + (long long)n * (16 * KS) + 2 * j); s1 = __ldg(scale + (long long)n * (16 * KS) + 2 * j + 1);
                  h0 = __ldg(shift + (long long)n * (16 * KS) + 2 * j); h1 = __ldg(shift + (long long)n * (16 * KS) + 2 * j + 1);
                }
                e.v[0] = from_f<T>(xl_silu<FUSE>(fmaf(to_f<T>(e.v[0]), s0, h0)));
                e.v[1] = from_f<T>(xl_silu<FUSE>(fmaf(to_f<T>(e.v[1]), s1, h1)));
                v[j] = *reinterpret_cast<uint32_t*>(&e);
              }
              if (a_out != nullptr && i >= 1 && i <= BY && pz >= z0 && pz < zhi) {
                char* dst = const_cast<char*>(ab) + (long long)n * p.asn_b + (long long)pz * p.asd_b +
                            (long long)(y0 - 1 + i) * p.ash_b + (long long)xv * (32 * KS);
#pragma unroll
                for (int j = 0; j < W / 4; ++j)
                  *reinterpret_cast<uint4*>(dst + 16 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              }
            }
            // neighbours in x: lane - 1 / lane + 1, across warps through a 32-byte exchange, zeros at the line ends
            const uint32_t ex = sm_ex + par * (8u * 4u * W) ;
            if (lane == 0 || lane == 31) {
              const uint32_t dst = ex + (uint32_t)(q * 2 + (lane == 31 ? 1 : 0)) * (4u * W);
#pragma unroll
              for (int j = 0; j < W / 4; ++j) st_shared_v4(dst + 16u * j, make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
            }
#pragma unroll
            for (int j = 0; j < W; ++j) {
              lf[j] = __shfl_up_sync(0xffffffffu, v[j], 1);
              rt[j] = __shfl_down_sync(0xffffffffu, v[j], 1);
            }
            asm volatile("bar.sync 2, 128;" ::: "memory");
            if (lane == 0) {
              if (q == 0) {
#pragma unroll
                for (int j = 0; j < W; ++j) lf[j] = 0u;
              } else {
                const uint32_t s2 = ex + (uint32_t)((q - 1) * 2 + 1) * (4u * W);
#pragma unroll
                for (int j = 0; j < W / 4; ++j)
                  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                               : "=r"(lf[4 * j]), "=r"(lf[4 * j + 1]), "=r"(lf[4 * j + 2]), "=r"(lf[4 * j + 3])
                               : "r"(s2 + 16u * j));
              }
            }
            if (lane == 31) {
              if (q == 3) {
#pragma unroll
                for (int j = 0; j < W; ++j) rt[j] = 0u;
              } else {
                const uint32_t s2 = ex + (uint32_t)((q + 1) * 2) * (4u * W);
#pragma unroll
                for (int j = 0; j < W / 4; ++j)
                  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                               : "=r"(rt[4 * j]), "=r"(rt[4 * j + 1]), "=r"(rt[4 * j + 2]), "=r"(rt[4 * j + 3])
                               : "r"(s2 + 16u * j));
              }
            }
            par ^= 1u;
            mbar_wait(a_free + 8 * slot, aph ^ 1);
            tc_fence_after();
            const uint32_t ta = t_lane + (uint32_t)slot * ACOLS;
#pragma unroll
            for (int k = 0; k < KS; ++k) {
              tmem_st8(ta + (uint32_t)(0 * KS + k) * 8u, lf + 8 * k);
              tmem_st8(ta + (uint32_t)(1 * KS + k) * 8u, v + 8 * k);
              tmem_st8(ta + (uint32_t)(2 * KS + k) * 8u, rt + 8 * k);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full + 8 * slot);
            if (++slot == NA) { slot = 0; aph ^= 1; }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(raw_free + 8 * rs);
          if (++rs == NR) { rs = 0; rph ^= 1; }
        }
      }
    }
  } else {
    // ===================================================================== epilogue
    const int q = warp;
    const int xv = q * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (uint32_t c = 0; c < ACC; c += 16) tmem_st16_zero(t_lane + c);
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0)
      for (int o = 0; o < BY; ++o) mbar_arrive(acc_free + 8 * o);
    float bs[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) bs[c] = bias ? __ldg(bias + c) : 0.f;
    uint32_t pc = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      int n, z0, zhi, y0;
      decode(u, n, z0, zhi, y0);
      float s1[16], s2[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) s1[c] = s2[c] = 0.f;
      for (int pz = z0 - 1; pz <= zhi; ++pz, ++pc) {
        const int zo = pz - 1;
        const bool sv = zo >= z0 && zo < zhi;
        const uint32_t cs = (uint32_t)xl_mod3(zo) * 16u;
        const bool last = pz == zhi;
#pragma unroll 1
        for (int o = 0; o < BY; ++o) {
          mbar_wait(acc_full + 8 * o, pc & 1);
          tc_fence_after();
          const uint32_t col = t_lane + (uint32_t)o * 48u;
          const bool st = sv && y0 + o < p.h;
          uint32_t r[16];
          if (st) {
            tmem_ld16(col + cs, r);
            tmem_ld_wait();
          }
          if (last) {
            tmem_st16_zero(col);
            tmem_st16_zero(col + 16u);
            tmem_st16_zero(col + 32u);
          } else {
            tmem_st16_zero(col + cs);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_free + 8 * o);
          if (st) {
            T* yp = y + (long long)n * p.ysn + (long long)zo * p.ysd + (long long)(y0 + o) * p.ysh + (long long)xv * p.ysw;
            float f[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) f[c] = __uint_as_float(r[c]) + bs[c];
            if (p.accumulate) {
              const Pack<T, 8> o0 = *reinterpret_cast<const Pack<T, 8>*>(yp);
              const Pack<T, 8> o1 = *reinterpret_cast<const Pack<T, 8>*>(yp + 8);
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                f[c] += to_f<T>(o0.v[c]);
                f[8 + c] += to_f<T>(o1.v[c]);
              }
            }
            Pack<T, 8> w0, w1;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              w0.v[c] = from_f<T>(f[c]);
              w1.v[c] = from_f<T>(f[8 + c]);
            }
            *reinterpret_cast<Pack<T, 8>*>(yp) = w0;
            *reinterpret_cast<Pack<T, 8>*>(yp + 8) = w1;
            if (stats != nullptr) {
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                const float vv = to_f<T>(c < 8 ? w0.v[c] : w1.v[c - 8]);
                s1[c] += vv;
                s2[c] = fmaf(vv, vv, s2[c]);
              }
            }
          }
        }
      }
      if (stats != nullptr) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float a = warp_sum(s1[c]), b = warp_sum(s2[c]);
          if (lane == 0) {
            atomicAdd(stats + ((long long)n * 16 + c) * 2, (double)a);
            atomicAdd(stats + ((long long)n * 16 + c) * 2 + 1, (double)b);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}

// Weights of the x-line kernel: [r][dy][dx][k][row = s*16 + co][kk], ci = 16 k + kk, tap dz = (r + 1 - s) mod 3, each
// (48 x 16) tile in the SWIZZLE_32B K-major shared-memory image (16-byte half kk >> 3 of row rr sits at half ^ ((rr >> 2) & 1)),
// so the whole matrix is ONE linear bulk copy.  flip: dgrad operand W'[ci][co][2-dz][2-dy][2-dx].
template <typename T>
__global__ void pack_weight_xline_kernel(const float* __restrict__ w, T* __restrict__ out, int cout, int cin, int flip) {
  const int CO = flip ? cin : cout, CI = flip ? cout : cin;   // CO == 16
  const int ks = CI / 16;
  const int total = 27 * ks * 48 * 16;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int t = idx;
    const int kk = t % 16; t /= 16;
    const int rr = t % 48; t /= 48;
    const int k = t % ks; t /= ks;
    const int dx = t % 3; t /= 3;
    const int dy = t % 3; t /= 3;
    const int r = t;
    const int s = rr / 16, co = rr % 16;
    const int dz = ((r + 1 - s) % 3 + 3) % 3;
    const int ci = k * 16 + kk;
    float v;
    if (flip) v = w[((((int64_t)ci * cin + co) * 3 + (2 - dz)) * 3 + (2 - dy)) * 3 + (2 - dx)];
    else v = w[((((int64_t)co * cin + ci) * 3 + dz) * 3 + dy) * 3 + dx];
    (void)CO;
    const int tile = ((r * 3 + dy) * 3 + dx) * ks + k;
    const int off = rr * 16 + ((((kk >> 3) ^ ((rr >> 2) & 1))) << 3) + (kk & 7);
    out[(int64_t)tile * (48 * 16) + off] = from_f<T>(v);
  }
}

// One MMA with A in tensor memory against exact integers: D[128][48] = A[128][16] * B[48][16]^T, A written with tcgen05.st
// (thread = lane = row, column j = elements 2j, 2j+1), B in the SWIZZLE_32B image of pack_weight_xline_kernel.
__global__ void __launch_bounds__(128, 1) xline_selftest_kernel(float* __restrict__ out, uint32_t idesc) {
  __shared__ __align__(1024) __nv_bfloat16 s_b[48 * 16];
  __shared__ uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
  if (t == 0) { mbar_init(smem_u32(&s_bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 128);
  for (int idx = t; idx < 48 * 16; idx += 128) {
    const int kk = idx % 16, rr = idx / 16;
    const int off = rr * 16 + ((((kk >> 3) ^ ((rr >> 2) & 1))) << 3) + (kk & 7);
    s_b[off] = __float2bfloat16_rn((float)((rr * 3 + kk * 5) % 7 - 3));
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  uint32_t a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    Pack<__nv_bfloat16, 2> e;
    e.v[0] = __float2bfloat16_rn((float)((t + 2 * j) % 5 - 2));
    e.v[1] = __float2bfloat16_rn((float)((t * 2 + 2 * j + 1) % 9 - 4));
    a[j] = *reinterpret_cast<uint32_t*>(&e);
  }
  tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + 64u, a);
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0 && elect_one()) {
    const uint32_t b_hi = (256u >> 4) | (1u << 14) | ((uint32_t)kSwizzle32 << 29);
    const uint32_t b_lo = ((smem_u32(s_b) >> 4) & 0x3FFFu) | (1u << 16);
    umma_f16_ts(tmem, tmem + 64u, b_lo, b_hi, idesc, 0u);
    umma_commit(smem_u32(&s_bar));
  }
  mbar_wait(smem_u32(&s_bar), 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 48; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 16; ++c) out[t * 48 + c0 + c] = __uint_as_float(r[c]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

bool conv_xline_ok(const ActView& x, const ActView& y, int kd, int kh, int kw) {
  if (kd != 3 || kh != 3 || kw != 3) return false;
  if (x.dtype != y.dtype || (x.dtype != B200_BF16 && x.dtype != B200_F16)) return false;
  if (x.w != 128 || y.c != 16 || (x.c != 16 && x.c != 48)) return false;
  if (x.n != y.n || x.d != y.d || x.h != y.h || x.w != y.w) return false;
  if (x.sw != x.c || x.sh != (int64_t)x.w * x.c) return false;                  // dense lines, contiguous lines in a plane
  if (((uintptr_t)x.data & 15) || (x.sd * 2) % 16 || (x.sn * 2) % 16) return false;
  if (((uintptr_t)y.data & 15) || (y.sw * 2) % 16 || (y.sh * 2) % 16 || (y.sd * 2) % 16 || (y.sn * 2) % 16) return false;
  return true;
}

static int xline_mode() {
  static int m = -1;
  if (m < 0) {
    const char* e = getenv("B200_XLINE");
    m = e ? atoi(e) : 1;
  }
  return m;
}
bool conv_xline_enabled() { return xline_mode() != 0; }

template <typename T, int KS, int BY>
static int launch_xline(const ActView& x, const void* w, const float* bias, const ActView& y, const ActView* a_out, const float* scale,
                        const float* shift, int fuse, double* stats, XlineParams p, cudaStream_t st) {
  constexpr int NR = KS == 1 ? 8 : 3;
  const size_t smem = 27u * KS * kXlTileBytes + (size_t)NR * 2u * 4096u * KS + 2u * 8u * 32u * KS + 1024u;
  p.bands = (int)ceil_div(x.h, BY);
  {  // z chunks: whole waves of units over the SMs against the two halo planes every chunk re-reads
    double best = -1.0;
    const int kmax = x.d < 32 ? x.d : 32;
    for (int k = 1; k <= kmax; ++k) {
      const int zc = (int)ceil_div(x.d, k);
      const int kk = (int)ceil_div(x.d, zc);
      const int64_t units = (int64_t)x.n * p.bands * kk;
      const int64_t waves = ceil_div(units, sm_count());
      const double eff = (double)units / (double)(waves * sm_count()) * (double)zc / (double)(zc + 2);
      if (eff > best + 1e-9) { best = eff; p.zc = zc; p.zchunks = kk; }
    }
  }
  p.units = x.n * p.bands * p.zchunks;
  const int grid = p.units < sm_count() ? p.units : sm_count();
  T* ap = a_out ? (T*)a_out->data : nullptr;
#define XL_LAUNCH(F)                                                                                                   \
  {                                                                                                                    \
    auto kern = conv_fprop_xline_kernel<T, KS, BY, F>;                                                                 \
    B200_CUDA(raise_dyn_smem_cap(kern));                                                                               \
    kern<<<grid, 320, smem, st>>>((const T*)x.data, (const T*)w, bias, (T*)y.data, ap, scale, shift, stats, p);       \
  }
  if (fuse == 0) XL_LAUNCH(0)
  else if (fuse == 1) XL_LAUNCH(1)
  else XL_LAUNCH(2)
#undef XL_LAUNCH
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// fuse: 0 = the input is used as it is; 1 / 2 = a = silu(x * scale[n,c] + shift[n,c]) in front of the convolution (exact /
// one-MUFU chain), a_out (optional, dense, same shape as x) receives a.
int conv_fprop_xline_v(const ActView& x, const void* w, const float* bias, const ActView& y, int accumulate, const ActView* a_out,
                       const float* scale, const float* shift, int fuse, double* stats, cudaStream_t st) {
  B200_CHECK_ARG(conv_xline_ok(x, y, 3, 3, 3), "conv_fprop(xline): unsupported operands");
  B200_CHECK_ARG(fuse == 0 || (scale && shift), "conv_fprop(xline): fused normalisation needs scale and shift");
  B200_CHECK_ARG(fuse >= 0 && fuse <= 2, "conv_fprop(xline): bad fuse mode");
  if (a_out) {
    B200_CHECK_ARG(fuse != 0, "conv_fprop(xline): a_out without a fused activation");
    B200_CHECK_ARG(a_out->dtype == x.dtype && a_out->n == x.n && a_out->d == x.d && a_out->h == x.h && a_out->w == x.w &&
                   a_out->c == x.c && a_out->sw == x.c && a_out->sh == (int64_t)x.w * x.c && !((uintptr_t)a_out->data & 15) &&
                   (a_out->sd * 2) % 16 == 0 && (a_out->sn * 2) % 16 == 0, "conv_fprop(xline): a_out must be a dense copy of x's shape");
  }
  XlineParams p{};
  p.n = x.n; p.d = x.d; p.h = x.h;
  p.xsh_b = x.sh * 2; p.xsd_b = x.sd * 2; p.xsn_b = x.sn * 2;
  if (a_out) { p.ash_b = a_out->sh * 2; p.asd_b = a_out->sd * 2; p.asn_b = a_out->sn * 2; }
  p.ysw = y.sw; p.ysh = y.sh; p.ysd = y.sd; p.ysn = y.sn;
  p.accumulate = accumulate;
  p.idesc = make_idesc(x.dtype == B200_BF16, 48, 0, 0);
  if (x.dtype == B200_BF16) {
    if (x.c == 16) return launch_xline<__nv_bfloat16, 1, 8>(x, w, bias, y, a_out, scale, shift, fuse, stats, p, st);
    return launch_xline<__nv_bfloat16, 3, 4>(x, w, bias, y, a_out, scale, shift, fuse, stats, p, st);
  }
  if (x.c == 16) return launch_xline<__half, 1, 8>(x, w, bias, y, a_out, scale, shift, fuse, stats, p, st);
  return launch_xline<__half, 3, 4>(x, w, bias, y, a_out, scale, shift, fuse, stats, p, st);
}

}  // namespace sm100
}  // namespace b200

using namespace b200;

B200_EXPORT int b200_xline_selftest(double* max_err, int32_t verbose, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  float* d_out = nullptr;
  B200_CUDA(cudaMalloc(&d_out, sizeof(float) * 128 * 48));
  sm100::xline_selftest_kernel<<<1, 128, 0, st>>>(d_out, sm100::make_idesc(1, 48, 0, 0));
  B200_LAUNCH_CHECK();
  std::vector<float> h(128 * 48);
  B200_CUDA(cudaMemcpyAsync(h.data(), d_out, sizeof(float) * h.size(), cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_out);
  double worst = 0.0;
  for (int t = 0; t < 128; ++t)
    for (int rr = 0; rr < 48; ++rr) {
      double ref = 0.0;
      for (int kk = 0; kk < 16; ++kk) {
        const int j = kk / 2;
        const double a = (kk & 1) ? (double)((t * 2 + 2 * j + 1) % 9 - 4) : (double)((t + 2 * j) % 5 - 2);
        ref += a * (double)((rr * 3 + kk * 5) % 7 - 3);
      }
      const double e = fabs(ref - (double)h[t * 48 + rr]);
      if (e > worst) worst = e;
      if (verbose > 1 && e > 0.5 && t < 4) printf("xline_selftest: row %d col %d got %g expected %g\n", t, rr, h[t * 48 + rr], ref);
    }
  if (verbose) { printf("xline_selftest: max abs error %g\n", worst); fflush(stdout); }
  if (max_err) *max_err = worst;
  return B200_OK;
}

B200_EXPORT int b200_conv_xline_supported(const b200_tensor* x, const b200_tensor* y, int32_t kd, int32_t kh, int32_t kw) {
  if (!x || !y || !x->data || !y->data) return 0;
  return sm100::conv_xline_enabled() && sm100::conv_xline_ok(sm100::view_of(x), sm100::view_of(y), kd, kh, kw) ? 1 : 0;
}

B200_EXPORT int b200_pack_conv_weight_xline(const float* w, void* packed, int32_t dtype, int32_t cout, int32_t cin,
                                            int32_t flip_transpose, void* stream) {
  B200_CHECK_ARG(w && packed, "pack_conv_weight_xline: null pointer");
  const int CO = flip_transpose ? cin : cout, CI = flip_transpose ? cout : cin;
  B200_CHECK_ARG(CO == 16 && (CI == 16 || CI == 48), "pack_conv_weight_xline: (Cout', Cin') = (%d, %d) not supported", CO, CI);
  cudaStream_t st = (cudaStream_t)stream;
  const int total = 27 * (CI / 16) * 48 * 16;
  const unsigned blocks = (unsigned)ceil_div(total, 256);
  if (dtype == B200_BF16)
    sm100::pack_weight_xline_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(w, (__nv_bfloat16*)packed, cout, cin, flip_transpose);
  else if (dtype == B200_F16)
    sm100::pack_weight_xline_kernel<__half><<<blocks, 256, 0, st>>>(w, (__half*)packed, cout, cin, flip_transpose);
  else {
    set_error("pack_conv_weight_xline: 16-bit dtypes only");
    return B200_ERR_ARG;
  }
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_conv_fprop_xline(const b200_tensor* x, const void* w_packed, const float* bias, const b200_tensor* y,
                                      int32_t accumulate, const float* scale, const float* shift, int32_t fuse,
                                      const b200_tensor* a_out, double* sums, void* stream) {
  B200_CHECK_ARG(x && y && w_packed, "conv_fprop_xline: null pointer");
  B200_CHECK_ARG(check_tensor(x, "conv_fprop_xline.x") && check_tensor(y, "conv_fprop_xline.y"), "%s", b200_last_error());
  B200_CHECK_ARG(!a_out || check_tensor(a_out, "conv_fprop_xline.a_out"), "%s", b200_last_error());
  sm100::ActView av{};
  if (a_out) av = sm100::view_of(a_out);
  return sm100::conv_fprop_xline_v(sm100::view_of(x), w_packed, bias, sm100::view_of(y), accumulate, a_out ? &av : nullptr, scale,
                                   shift, fuse, sums, (cudaStream_t)stream);
}
