// sm_100a primitives used by the tensor-core kernels: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma /
// commit / ld) as inline PTX, plus host helpers that build CUtensorMap descriptors through the driver entry point
// (no link-time dependency on libcuda).
#pragma once
#include "common.cuh"

#include <cuda.h>

namespace b200 {
namespace sm100 {

// ------------------------------------------------------------------------------------------------ device side
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (CUDA error) after ~2 s instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("biapy_b200: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3,
                                            int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"((uint64_t)tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)tmap), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// shared -> global through a rank-5 tensor map (bulk async-group completion): plain store and element-wise add.  Boxes that
// stick out of the tensor are clipped by the TMA unit.
__device__ __forceinline__ void tma_store_5d(const void* tmap, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"((uint64_t)tmap), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_5d(const void* tmap, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.reduce.async.bulk.tensor.5d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"((uint64_t)tmap), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tmap) : "memory");
}

// 16-byte asynchronous global->shared copy; src_bytes == 0 zero-fills the destination ('same' padding / tile tails)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// arrive on `bar` once all cp.async issued so far by this thread have landed (counted in the barrier's init count)
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; kind::f16 covers fp16 and bf16 inputs with fp32 accumulation
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// One lane of a converged warp (the compiler knows the guarded region is single-threaded, so tcgen05 operands move to
// uniform registers without a per-lane waterfall loop)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// Keeps a loop constant in a register: without it the compiler re-loads kernel parameters from the constant bank
// after every asm volatile("...": "memory"), and the single MMA-issuing thread stalls on each of those loads.
__device__ __forceinline__ uint32_t in_reg(uint32_t v) {
  asm volatile("" : "+r"(v));
  return v;
}
__device__ __forceinline__ int in_reg(int v) {
  asm volatile("" : "+r"(v));
  return v;
}

// Descriptor passed as (low word, high word): only the low word (start address, LBO) changes between K steps, slots and
// taps, so stepping a descriptor is ONE 32-bit add.  The issuing thread runs ~8 cycles per dependent scalar instruction
// (B200_DBG counters, profiles/README.md): at N = 64 the 64-bit descriptor arithmetic cost more than the MMA itself.
__device__ __forceinline__ void umma_f16_split(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <int KS>
__device__ __forceinline__ void umma_ksteps_split(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                  uint32_t astep16, uint32_t bstep16, uint32_t idesc, uint32_t acc_first) {
#pragma unroll
  for (int k = 0; k < KS; ++k)
    umma_f16_split(tmem_d, a_lo + astep16 * k, a_hi, b_lo + bstep16 * k, b_hi, idesc, k == 0 ? acc_first : 1u);
}

// KS consecutive K = 16 steps of one operand pair.  `adesc` / `bdesc` are complete descriptors of the first step; a step
// advances the 14-bit start-address field by 32 bytes (>> 4 = 2), which never carries out of the field for shared
// memory addresses.  ncu on the first kernels showed ~200 cycles of dependent scalar work per MMA when descriptors
// were rebuilt per step (profiles/README.md); this form is two 64-bit adds per MMA.
template <int KS>
__device__ __forceinline__ void umma_ksteps(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc_first) {
#pragma unroll
  for (int k = 0; k < KS; ++k) umma_f16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k == 0 ? acc_first : 1u);
}

// Same with explicit per-step advances (in 16-byte units) for MN-major operands whose K steps are whole swizzle-atom rows.
template <int KS>
__device__ __forceinline__ void umma_ksteps_strided(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t astep16,
                                                    uint32_t bstep16, uint32_t idesc, uint32_t acc_first) {
#pragma unroll
  for (int k = 0; k < KS; ++k)
    umma_f16(tmem_d, adesc + (uint64_t)(astep16 * k), bdesc + (uint64_t)(bstep16 * k), idesc, k == 0 ? acc_first : 1u);
}

// 32 lanes x 16 consecutive fp32 columns: thread t of the warp receives row (lane base + t)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, SM100 version field = 1)
//   bits [0,14)  start address >> 4      bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4 bits [46,48) version (1)     bits [61,64) layout type
enum : uint32_t { kSwizzleNone = 0, kSwizzle128 = 2, kSwizzle64 = 4, kSwizzle32 = 6 };
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}

// Instruction descriptor for kind::f16 (cute::UMMA::InstrDescriptor): fp32 accumulate, M = 128
//   c_format [4,6) = 1 (F32); a_format [7,10), b_format [10,13): 0 = F16, 1 = BF16; a_major bit 15, b_major bit 16
//   (0 = K-major, 1 = MN-major); n_dim [17,23) = N >> 3; m_dim [24,29) = M >> 4
inline uint32_t make_idesc(int is_bf16, int n, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= (uint32_t)(is_bf16 ? 1 : 0) << 7;
  d |= (uint32_t)(is_bf16 ? 1 : 0) << 10;
  d |= (uint32_t)(a_mn_major ? 1 : 0) << 15;
  d |= (uint32_t)(b_mn_major ? 1 : 0) << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(128 >> 4) << 24;
  return d;
}

// -------------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn();   // resolved once via cudaGetDriverEntryPoint

inline CUtensorMapSwizzle swizzle_for_bytes(int inner_bytes) {
  return inner_bytes >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (inner_bytes >= 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// Channels-last activation view with explicit element strides per spatial axis; lets a stride-s sub-lattice of a
// tensor (the s^3 output phases of a transposed convolution) be addressed as an ordinary tensor by TMA and epilogues.
struct ActView {
  void* data;
  int dtype;
  int n, d, h, w, c;
  int64_t sw, sh, sd, sn;   // element strides of x, y, z, batch
};
inline ActView view_of(const b200_tensor* t) {
  return ActView{t->data, t->dtype, t->n, t->d, t->h, t->w, t->c, t->ld, (int64_t)t->w * t->ld, (int64_t)t->h * t->w * t->ld,
                 (int64_t)t->d * t->h * t->w * t->ld};
}
// phase (a, b, c) of the (sd, sh, sw)-strided sub-lattice of `fine`: dims = fine dims / stride
inline ActView phase_view(const b200_tensor* fine, int sd, int sh, int sw, int a, int b, int c) {
  ActView v = view_of(fine);
  v.data = (char*)fine->data + (((int64_t)a * fine->h + b) * fine->w + c) * fine->ld * 2;
  v.d = fine->d / sd; v.h = fine->h / sh; v.w = fine->w / sw;
  v.sw *= sw; v.sh *= sh; v.sd *= sd;
  return v;
}
// rank-5 map (C, W, H, D, N), box (ck, bw, bh, bd, 1)
int make_act_tmap(CUtensorMap* out, const ActView& t, int ck, int bw, int bh, int bd);
// row-major matrix [rows][cols] of 16-bit elements: rank-2 map (cols, rows), box (box_cols, box_rows)
int make_matrix_tmap(CUtensorMap* out, const void* base, int dtype, int64_t rows, int64_t cols, int box_rows, int box_cols);

}  // namespace sm100
}  // namespace b200
