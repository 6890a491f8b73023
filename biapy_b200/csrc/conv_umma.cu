// tcgen05 / TMA implicit-GEMM convolution for sm_100a (bf16 / fp16 storage, fp32 accumulation in TMEM).
//
//   fprop / dgrad:  D[128 voxels][Cout tile] = sum over (tap, Cin chunk) A_tap[128 voxels][CK] * W[Cout tile][tap, CK]
//     A_tap is one TMA box of the channels-last activation tensor shifted by the tap offset; out-of-bounds voxels are
//     zero-filled by the TMA unit, which IS the 'same' padding -- no im2col buffer, no halo code.  Weights
//     [Cout][tap][Cin] are a plain K-major matrix.  Both land in 32/64/128-byte-swizzled shared memory and feed
//     tcgen05.mma directly (M = 128, N = Cout tile <= 256, K = 16 per instruction).
//   Warp roles (192 threads, persistent over tiles): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner,
//     warps 2-5 = epilogue (TMEM -> registers -> +bias (+old value) -> 16-bit -> global, 32 B per thread per 16 columns).
//   Two TMEM accumulator buffers let the epilogue of tile i overlap the main loop of tile i+1.
#include "umma.cuh"

#include <mutex>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <string>

namespace b200 {
int conv_bias_grad(const b200_tensor* dy, float* dbias, cudaStream_t st);   // conv_simt.cu

namespace sm100 {

// ---------------------------------------------------------------------------------------- tensor-map helpers
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

static CUtensorMapDataType tm_dtype(int dt) {
  return dt == B200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
}

int make_act_tmap(CUtensorMap* out, const ActView& t, int ck, int bw, int bh, int bd) {
  EncodeTiledFn fn = encode_tiled_fn();
  B200_CHECK_ARG(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  const uint64_t es = 2;
  cuuint64_t dims[5] = {(cuuint64_t)t.c, (cuuint64_t)t.w, (cuuint64_t)t.h, (cuuint64_t)t.d, (cuuint64_t)t.n};
  cuuint64_t strides[4] = {(cuuint64_t)t.sw * es, (cuuint64_t)t.sh * es, (cuuint64_t)t.sd * es, (cuuint64_t)t.sn * es};
  cuuint32_t box[5] = {(cuuint32_t)ck, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bd, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(out, tm_dtype(t.dtype), 5, t.data, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_for_bytes(ck * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation) failed with %d (c=%d sw=%lld box=%d,%d,%d,%d)", (int)r,
                 t.c, (long long)t.sw, ck, bw, bh, bd);
  return B200_OK;
}

int make_matrix_tmap(CUtensorMap* out, const void* base, int dtype, int64_t rows, int64_t cols, int box_rows, int box_cols) {
  EncodeTiledFn fn = encode_tiled_fn();
  B200_CHECK_ARG(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, tm_dtype(dtype), 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_for_bytes(box_cols * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(matrix) failed with %d (rows=%lld cols=%lld box=%d,%d)", (int)r,
                 (long long)rows, (long long)cols, box_rows, box_cols);
  return B200_OK;
}

// ------------------------------------------------------------------------------------------------- fprop kernel
struct FpropParams {
  int n, d, h, w, cin, cout;
  int kd, kh, kw;
  int bd, bh, bw;                       // voxel tile (bd*bh*bw == 128)
  int tiles_d, tiles_h, tiles_w, tiles_n;
  int num_tiles;
  int ck, chunks;                       // channels per K block, Cin / ck
  int nt;                               // Cout tile (multiple of 16, <= 256)
  int stages;
  uint32_t a_bytes, b_bytes, stage_bytes;
  uint32_t layout, sbo;                 // UMMA swizzle code and stride-byte-offset for this ck
  uint32_t idesc;
  uint32_t tmem_cols;
  int64_t ysw, ysh, ysd, ysn;           // output element strides (x, y, z, batch)
  int accumulate;
  // transposed-convolution mode: GEMM column c = phase * pcout + co is scattered to the fine voxel
  // (z*usd + a, y*ush + b, x*usw + c') of phase (a, b, c'); ys* are then the strides of the FINE tensor.  usd == 0: off.
  int usd, ush, usw, pcout;
  // phase-gather mode (dgrad of a transposed convolution): tap t reads its A box through PhaseMaps::m[t] at unshifted
  // coordinates and its weights from rows [t * wrows, t * wrows + cout) of the packed matrix; kd = phases, kh = kw = 1
  int phases, wrows;
  // transposed mode, TMA-store epilogue: PhaseMaps::m[t] describes output phase t of the fine tensor (box 16 channels x voxel
  // tile); a (128 voxels x 16 channels) block is staged at epi_off (2 x 4 KB) and leaves with one bulk tensor store
  int epi_tma;
  uint32_t epi_off;
  int epi_cbox;                         // channels per output box: 64 / 32 / 16 (SWIZZLE_128B / 64B / none)
};

constexpr int kMaxStages = 12;

// Tensor maps of the (up to 8) output phases of a stride-s transposed convolution: phase t of the fine tensor is the
// strided sub-lattice (a, b, c) seen as an ordinary coarse tensor (phase_view).  Lets ONE launch run over all phases:
// K blocks of the dgrad GEMM / N boxes of the wgrad GEMM pick their map by phase instead of one launch per phase.
struct alignas(64) PhaseMaps { CUtensorMap m[8]; };

template <typename T>
__global__ void __launch_bounds__(224, 1)
conv_fprop_umma_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                       const __grid_constant__ PhaseMaps pm, const float* __restrict__ bias, T* __restrict__ y,
                       const FpropParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_bar[2 * kMaxStages + 4];
  __shared__ uint32_t s_tmem;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = smem_u32(&s_bar[0]);
  const uint32_t bar_empty = smem_u32(&s_bar[kMaxStages]);
  const uint32_t bar_tfull = smem_u32(&s_bar[2 * kMaxStages]);
  const uint32_t bar_tempty = smem_u32(&s_bar[2 * kMaxStages + 2]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 2);                    // activation producer + weight producer
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, 128);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1) tmem_alloc(smem_u32(&s_tmem), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  const int taps = p.kd * p.kh * p.kw;
  const int num_kb = taps * p.chunks;
  const bool pmode = p.phases > 0;
  const int pd = pmode ? 0 : p.kd / 2, ph = p.kh / 2, pw = p.kw / 2;

  auto decode = [&](int tile, int& n, int& z0, int& y0, int& x0, int& n0) {
    int t = tile;
    n0 = (t % p.tiles_n) * p.nt; t /= p.tiles_n;
    x0 = (t % p.tiles_w) * p.bw; t /= p.tiles_w;
    y0 = (t % p.tiles_h) * p.bh; t /= p.tiles_h;
    z0 = (t % p.tiles_d) * p.bd; t /= p.tiles_d;
    n = t;
  };

  if (warp == 0 || warp == 6) {
    // =================================================================== TMA producers
    // warp 0: activation boxes, warp 6: weight boxes.  One thread sustains one bulk-tensor copy per ~450 cycles whatever
    // its size (tools/pipe_rates.py), so the two streams are issued by different warps; both arrive on the stage's full
    // barrier (count 2) with their own byte counts.  Nested tap loops: no integer division in the issue loop.
    if (elect_one()) {
      const bool a_role = warp == 0;
      const int kd = p.kd, kh = p.kh, kw = p.kw, chunks = p.chunks, ck = p.ck, stages = p.stages;
      const uint32_t a_bytes = p.a_bytes, b_bytes = p.b_bytes, stage_bytes = p.stage_bytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int n, z0, y0, x0, n0;
        decode(tile, n, z0, y0, x0, n0);
        int kcol = 0;
        for (int dz = 0; dz < kd; ++dz)
          for (int dy = 0; dy < kh; ++dy)
            for (int dx = 0; dx < kw; ++dx)
              for (int ch = 0; ch < chunks; ++ch, kcol += ck) {
                mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                const uint32_t a_dst = smem0 + stage * stage_bytes;
                const uint32_t fb = bar_full + 8 * stage;
                if (a_role) {
                  mbar_expect_tx(fb, a_bytes);
                  if (pmode) tma_load_5d(a_dst, &pm.m[dz], fb, ch * ck, x0, y0, z0, n);
                  else tma_load_5d(a_dst, &tmap_x, fb, ch * ck, x0 + dx - pw, y0 + dy - ph, z0 + dz - pd, n);
                } else {
                  mbar_expect_tx(fb, b_bytes);
                  if (pmode) tma_load_2d(a_dst + a_bytes, &tmap_w, fb, ch * ck, dz * p.wrows + n0);
                  else tma_load_2d(a_dst + a_bytes, &tmap_w, fb, kcol, n0);
                }
                if (++stage == stages) { stage = 0; phase ^= 1; }
              }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ================================================================= MMA issuer
      // descriptor template + slot address in 16-byte units: two 64-bit adds per MMA (umma_ksteps)
      const int stages = p.stages, ksteps = p.ck / 16;
      const uint32_t idesc = p.idesc, stage16 = p.stage_bytes >> 4, aoff16 = p.a_bytes >> 4, s0 = smem0 >> 4;
      const uint64_t tmpl = make_smem_desc(0, 16, p.sbo, p.layout);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(bar_tempty + 8 * buf, ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem + buf * p.nt;
        uint32_t acc = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t a16 = s0 + (uint32_t)stage * stage16;
          const uint64_t ad = tmpl + a16, bd = tmpl + (a16 + aoff16);
          if (ksteps == 4) umma_ksteps<4>(d_tmem, ad, bd, idesc, acc);
          else if (ksteps == 2) umma_ksteps<2>(d_tmem, ad, bd, idesc, acc);
          else umma_ksteps<1>(d_tmem, ad, bd, idesc, acc);
          acc = 1;
          umma_commit(bar_empty + 8 * stage);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(bar_tfull + 8 * buf);
      }
    }
  } else {
    // =================================================================== epilogue (warps 2..5 -> TMEM lane quarter warp%4)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int lx = row % p.bw, ly = (row / p.bw) % p.bh, lz = row / (p.bw * p.bh);
    int it = 0;
    uint32_t up_box = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      int n, z0, y0, x0, n0;
      decode(tile, n, z0, y0, x0, n0);
      mbar_wait(bar_tfull + 8 * buf, (it >> 1) & 1);
      tc_fence_after();
      const int gz = z0 + lz, gy = y0 + ly, gx = x0 + lx;
      const bool valid = gz < p.d && gy < p.h && gx < p.w;
      const bool up = p.usd != 0;
      if (up && p.epi_tma) {
        // Transposed convolution, TMA-store epilogue.  The direct path below has every thread store the 32-byte piece of its
        // own coarse voxel: 32 pieces a fine-voxel pitch (or more) apart per warp instruction = 32 LSU wavefronts, and the
        // kernel (K = Cin: a couple of MMAs per tile) is nothing but its epilogue.  Here the whole accumulator tile is staged
        // -- one dense 4 KB box (128 rows x 16 channels) per (phase, channel block) column chunk -- behind ONE pair of
        // barriers, and leaves with one bulk tensor store per chunk through the map of its phase (the strided sub-lattice of the
        // fine tensor seen as a coarse tensor; partial tiles are clipped by the TMA unit).  Two staging buffers alternate per
        // tile.  (A first version synchronised per chunk: 32 barriers + 16 proxy fences per tile, 0.557 -> 0.508 ms only.)
        const bool issuer = threadIdx.x == 64;
        const uint32_t sbuf = smem0 + p.epi_off + (up_box & 1u) * (uint32_t)p.nt * 256u;
        const int ph_t0 = n0 / p.pcout, ph_co0 = n0 - ph_t0 * p.pcout;
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * p.nt);
        // box = 128 rows x epi_cbox channels (row pitch = epi_cbox * 2 bytes) in the shared-memory layout of the map's swizzle:
        // 16-byte piece c of row r sits at piece c ^ (r & 7) (128-byte rows), c ^ ((r >> 1) & 3) (64-byte rows), c (32-byte rows)
        const uint32_t cps = (uint32_t)p.epi_cbox >> 4;          // 16-channel chunks per box: 4 / 2 / 1
        const uint32_t lg = cps == 4 ? 2u : (cps == 2 ? 1u : 0u);
        const uint32_t swz = cps == 4 ? ((uint32_t)row & 7u) : (cps == 2 ? (((uint32_t)row >> 1) & 3u) : 0u);
        const uint32_t rbase = sbuf + ((uint32_t)row << (5u + lg));          // row pitch = 32 << lg bytes
        if (issuer) bulk_wait_group_read<1>();                   // the stores that read this buffer two tiles ago are done with it
        asm volatile("bar.sync 1, 128;" ::: "memory");
        int ph_co = ph_co0;
        for (int j0 = 0; j0 < p.nt && n0 + j0 < p.cout; j0 += 16) {
          uint32_t r[16];
          tmem_ld16(taddr + j0, r);
          tmem_ld_wait();
          const float* brow = bias ? bias + ph_co : nullptr;
          Pack<T, 8> w0, w1;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            w0.v[j] = from_f<T>(__uint_as_float(r[j]) + (brow ? __ldg(brow + j) : 0.f));
            w1.v[j] = from_f<T>(__uint_as_float(r[8 + j]) + (brow ? __ldg(brow + 8 + j) : 0.f));
          }
          const uint32_t chunk = (uint32_t)j0 >> 4;
          const uint32_t k2 = (chunk & (cps - 1u)) << 1;                      // first 16-byte piece of this chunk inside its row
          const uint32_t dst = rbase + ((chunk >> lg) << (12u + lg));         // box = 128 rows x (32 << lg) bytes
          st_shared_v4(dst + ((k2 ^ swz) << 4), *reinterpret_cast<const uint4*>(&w0));
          st_shared_v4(dst + (((k2 + 1u) ^ swz) << 4), *reinterpret_cast<const uint4*>(&w1));
          ph_co += 16;
          if (ph_co == p.pcout) ph_co = 0;
        }
        tc_fence_before();
        mbar_arrive(bar_tempty + 8 * buf);                       // the accumulator has been read: the next tile may use it
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (issuer) {
          int ph_t = ph_t0;
          ph_co = ph_co0;
          for (int j0 = 0; j0 < p.nt && n0 + j0 < p.cout; j0 += p.epi_cbox) {
            tma_store_5d(&pm.m[ph_t], sbuf + (((uint32_t)j0 >> (4u + lg)) << (12u + lg)), ph_co, x0, y0, z0, n);
            ph_co += p.epi_cbox;
            if (ph_co == p.pcout) { ph_co = 0; ++ph_t; }
          }
          bulk_commit_group();
        }
        ++up_box;
        continue;
      }
      T* ybase = up ? y + (int64_t)n * p.ysn + (int64_t)gz * p.usd * p.ysd + (int64_t)gy * p.ush * p.ysh + (int64_t)gx * p.usw * p.ysw
                    : y + (int64_t)n * p.ysn + (int64_t)gz * p.ysd + (int64_t)gy * p.ysh + (int64_t)gx * p.ysw + n0;
      // transposed mode: running (phase, channel) of the current 16-column chunk
      int ph_t = up ? n0 / p.pcout : 0, ph_co = up ? n0 - ph_t * p.pcout : 0;
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * p.nt);
      for (int j0 = 0; j0 < p.nt; j0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + j0, r);
        tmem_ld_wait();
        if (valid && n0 + j0 < p.cout) {
          T* yrow;
          const float* brow;
          if (up) {
            const int pc = ph_t % p.usw, pb = (ph_t / p.usw) % p.ush, pa = ph_t / (p.usw * p.ush);
            yrow = ybase + (int64_t)pa * p.ysd + (int64_t)pb * p.ysh + (int64_t)pc * p.ysw + ph_co;
            brow = bias ? bias + ph_co : nullptr;
          } else {
            yrow = ybase + j0;
            brow = bias ? bias + n0 + j0 : nullptr;
          }
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(r[j]) + (brow ? __ldg(brow + j) : 0.f);
          if (p.accumulate) {
            Pack<T, 8> o0 = *reinterpret_cast<const Pack<T, 8>*>(yrow);
            Pack<T, 8> o1 = *reinterpret_cast<const Pack<T, 8>*>(yrow + 8);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              f[j] += to_f<T>(o0.v[j]);
              f[8 + j] += to_f<T>(o1.v[j]);
            }
          }
          Pack<T, 8> w0, w1;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            w0.v[j] = from_f<T>(f[j]);
            w1.v[j] = from_f<T>(f[8 + j]);
          }
          *reinterpret_cast<Pack<T, 8>*>(yrow) = w0;
          *reinterpret_cast<Pack<T, 8>*>(yrow + 8) = w1;
        }
        if (up) {
          ph_co += 16;
          if (ph_co == p.pcout) { ph_co = 0; ++ph_t; }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * buf);
    }
    if (p.epi_tma && threadIdx.x == 64) bulk_wait_group_read<0>();       // the staging buffers must outlive their stores
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------- x-folded fprop kernel
// For the small-channel layers (Cin <= 96 at 64^3..128^3, 60 % of the model FLOPs) one TMA box row of the direct kernel
// carries only 32..64 bytes and the TMA unit, not L2 or the tensor pipe, limits throughput (profiles/README.md).
// Folding 4 consecutive x-voxels into ONE GEMM row makes the rows fat and the instruction N large:
//     D[(y,z) row][ (j, co) ] = sum_{dz,dy} sum_{xi<6, ci} X[z+dz, y+dy, 4g-1+xi, ci] * Wt[(j,co)][(dz,dy,xi,ci)]
// with the block-Toeplitz weights Wt[(j,co)][(dz,dy,xi,ci)] = W[co][ci][dz][dy][xi-j] for 0 <= xi-j <= 2, else 0.
// The 6 input voxels of a row are 6*Cin contiguous elements (dense channels-last), fetched as 128-byte (64 el) and
// 64-byte (32 el) boxes of a rank-4 map (W*C, H, D, N): per output voxel the TMA moves 4.5 rows (Cin = 16) instead of 27,
// the MMA runs with N = 4*Cout (64..256) instead of 16..64, and each thread of the epilogue stores 4 voxels.
// Half of the issued MACs hit Toeplitz zeros; the tensor pipe is not the bottleneck of these layers.
struct XfoldParams {
  int n, d, h, w, cin, cout;
  int kd, kh, kw;                       // kw in {1, 3}; a GEMM row covers win = 3 + kw input voxels
  int bh, bd;                           // rows of a tile: bh * bd == 128 (y fastest)
  int groups_x, tiles_h, tiles_d, num_tiles;
  int kx;                               // win * cin
  int boxes64, has32;                   // K boxes per (dz, dy): boxes64 of 64 elements, then one of 32 if has32
  int nt;                               // 4 * cout
  int stages;
  uint32_t a_bytes, stage_bytes;        // A tile of a 64-element box; stage = A + B (max box)
  uint32_t idesc, tmem_cols;
  int64_t ysw, ysh, ysd, ysn;
  int accumulate;
};

template <typename T>
__global__ void __launch_bounds__(224, 1)
conv_fprop_xfold_kernel(const __grid_constant__ CUtensorMap tmx64, const __grid_constant__ CUtensorMap tmx32,
                        const __grid_constant__ CUtensorMap tmw64, const __grid_constant__ CUtensorMap tmw32,
                        const float* __restrict__ bias, T* __restrict__ y, const XfoldParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_bar[2 * kMaxStages + 4];
  __shared__ uint32_t s_tmem;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = smem_u32(&s_bar[0]);
  const uint32_t bar_empty = smem_u32(&s_bar[kMaxStages]);
  const uint32_t bar_tfull = smem_u32(&s_bar[2 * kMaxStages]);
  const uint32_t bar_tempty = smem_u32(&s_bar[2 * kMaxStages + 2]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 2);                    // activation producer + weight producer
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, 128);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmx64); tma_prefetch_desc(&tmx32); tma_prefetch_desc(&tmw64); tma_prefetch_desc(&tmw32);
  }
  if (warp == 1) tmem_alloc(smem_u32(&s_tmem), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const int pd = p.kd / 2, ph = p.kh / 2;
  const int nboxes = p.boxes64 + p.has32;

  auto decode = [&](int tile, int& n, int& z0, int& y0, int& g) {
    int t = tile;
    g = t % p.groups_x; t /= p.groups_x;
    y0 = (t % p.tiles_h) * p.bh; t /= p.tiles_h;
    z0 = (t % p.tiles_d) * p.bd; t /= p.tiles_d;
    n = t;
  };

  if (warp == 0 || warp == 6) {
    // =================================================================== TMA producers
    // warp 0 streams the activation boxes, warp 6 the weight boxes (one thread sustains only one bulk-tensor copy per
    // ~450 cycles, tools/pipe_rates.py); both arrive on the stage's full barrier (count 2) with their own byte counts
    if (elect_one()) {
      const bool a_role = warp == 0;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int n, z0, y0, g;
        decode(tile, n, z0, y0, g);
        const int e0 = (4 * g - p.kw / 2) * p.cin;       // first element of the input window (may be negative)
        int kcol = 0;
        for (int dz = 0; dz < p.kd; ++dz)
          for (int dy = 0; dy < p.kh; ++dy)
            for (int b = 0; b < nboxes; ++b) {
              const bool wide = b < p.boxes64;
              mbar_wait(bar_empty + 8 * stage, phase ^ 1);
              const uint32_t a_dst = smem0 + stage * p.stage_bytes;
              const uint32_t fb = bar_full + 8 * stage;
              const uint32_t wbytes = wide ? 128u : 64u;
              if (a_role) {
                mbar_expect_tx(fb, 128u * wbytes);
                asm volatile(
                    "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                    ::"r"(a_dst), "l"((uint64_t)(wide ? &tmx64 : &tmx32)), "r"(fb), "r"(e0 + b * 64), "r"(y0 + dy - ph),
                      "r"(z0 + dz - pd), "r"(n)
                    : "memory");
              } else {
                mbar_expect_tx(fb, (uint32_t)p.nt * wbytes);
                tma_load_2d(a_dst + p.a_bytes, wide ? &tmw64 : &tmw32, fb, kcol, 0);
              }
              kcol += wide ? 64 : 32;
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ================================================================= MMA issuer
      const int stages = p.stages, boxes64 = p.boxes64;
      const uint32_t idesc = p.idesc, stage16 = p.stage_bytes >> 4, aoff16 = p.a_bytes >> 4;
      const uint64_t tmpl128 = make_smem_desc(0, 16, 1024, kSwizzle128), tmpl64 = make_smem_desc(0, 16, 512, kSwizzle64);
      const uint32_t s0 = smem0 >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const int nkb = p.kd * p.kh * nboxes;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(bar_tempty + 8 * buf, ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem + buf * p.nt;
        int b = 0;
        uint32_t acc = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          const bool wide = b < boxes64;
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t a16 = s0 + (uint32_t)stage * stage16;
          if (wide) umma_ksteps<4>(d_tmem, tmpl128 + a16, tmpl128 + (a16 + aoff16), idesc, acc);
          else umma_ksteps<2>(d_tmem, tmpl64 + a16, tmpl64 + (a16 + aoff16), idesc, acc);
          acc = 1;
          umma_commit(bar_empty + 8 * stage);
          if (++stage == stages) { stage = 0; phase ^= 1; }
          if (++b == nboxes) b = 0;
        }
        umma_commit(bar_tfull + 8 * buf);
      }
    }
  } else {
    // =================================================================== epilogue
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int ly = row % p.bh, lz = row / p.bh;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      int n, z0, y0, g;
      decode(tile, n, z0, y0, g);
      mbar_wait(bar_tfull + 8 * buf, (it >> 1) & 1);
      tc_fence_after();
      const int gz = z0 + lz, gy = y0 + ly;
      const bool valid = gz < p.d && gy < p.h;
      T* ybase = y + (int64_t)n * p.ysn + (int64_t)gz * p.ysd + (int64_t)gy * p.ysh + (int64_t)(4 * g) * p.ysw;
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * p.nt);
      int j = 0, co = 0;                               // column = j * cout + co
      for (int c0 = 0; c0 < p.nt; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + c0, r);
        tmem_ld_wait();
        if (valid && 4 * g + j < p.w) {
          T* yrow = ybase + (int64_t)j * p.ysw + co;
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(r[i]) + (bias ? __ldg(bias + co + i) : 0.f);
          if (p.accumulate) {
            Pack<T, 8> o0 = *reinterpret_cast<const Pack<T, 8>*>(yrow);
            Pack<T, 8> o1 = *reinterpret_cast<const Pack<T, 8>*>(yrow + 8);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              f[i] += to_f<T>(o0.v[i]);
              f[8 + i] += to_f<T>(o1.v[i]);
            }
          }
          Pack<T, 8> w0, w1;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            w0.v[i] = from_f<T>(f[i]);
            w1.v[i] = from_f<T>(f[8 + i]);
          }
          *reinterpret_cast<Pack<T, 8>*>(yrow) = w0;
          *reinterpret_cast<Pack<T, 8>*>(yrow + 8) = w1;
        }
        co += 16;
        if (co == p.cout) { co = 0; ++j; }
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------- x-folded fprop, z-slab variant
// The x-folded kernel above re-reads every input row once per (dz, dy) tap: 9x for a 3x3x3 layer, and it re-reads the
// Toeplitz weights once per 128-row tile; it runs at the L2->SMEM limit (~8 TB/s, profiles/README.md).  This variant
// keeps the GEMM but changes what travels: for each dy one SLAB of (8*RT + kd - 1) z-lines x 16 y-rows is loaded once
// and the kd taps in z are served from it by moving the descriptor start address by whole 16-row lines (always a
// multiple of the 1 KB / 512 B swizzle atom, so the TMA-written swizzle stays valid); RT = 2 row tiles (16 z-lines)
// share every weight tile.  A bytes per voxel drop 2.4x, weight bytes 2x.  Separate rings for slabs and weight tiles.
// Optional per-(sample, channel) statistics fused into the x-slab epilogues (conv -> GroupNorm / InstanceNorm / BatchNorm,
// reference blocks.py:154-160): the epilogue holds every output element in registers, so
//   sums[n][c] += (sum v, sum v*v) of the stored (rounded) value v        (== b200_channel_sums of the output)
// comes out of the convolution and the normalisation's own read of the tensor for its statistics disappears.
// (The backward counterpart -- the (g, g*x) sums of the norm + activation backward inside the dgrad epilogue -- was built
// and measured: 0.69 ms against 0.28 + 0.29 ms for dgrad + stand-alone reduction on 16 -> 16 @128^3; it needs the norm input
// in the epilogue, i.e. a second operand stream, and was dropped.)
struct XslabParams {
  int n, d, h, w, cin, cout;
  int kd, kh, kw;
  int rt;                               // row tiles per CTA tile (1 or 2): 8*rt z-lines x 16 y-rows x 4 x-voxels
  int zl;                               // z-lines per slab = 8*rt + kd - 1
  int groups_x, tiles_h, tiles_d, num_tiles;
  int kx, boxes64, has32;               // kx = window elements per (dz, dy) incl. padding (xfold_geom)
  int xoff;                             // extra voxels at the window start (image-fed layers), 0 for Cin % 16 == 0
  int nt, nbuf;                         // 4*cout; TMEM accumulator buffers (1 or 2)
  int a_stages, b_stages;
  uint32_t a_bytes, b_bytes, b_off;     // ring slot sizes (max box) and offset of the weight ring
  uint32_t idesc, tmem_cols;
  int64_t ysw, ysh, ysd, ysn;
  int accumulate;
  int ablate;                           // diagnostics (B200_ABLATE): 1 = no TMA, 2 = no MMA, 4 = no epilogue stores
  long long* dbg;                       // diagnostics (B200_DBG): per-CTA cycle counters of the role loops, or nullptr
  double* stats;                        // fused channel sums [n][cout][2] (kernels instantiated with CB = cout / 16 > 0) or nullptr
  // TMA-store epilogue (xslab_epilogue_tma): byte offset of the 2 x 16 KB staging ring behind the operand rings, and the
  // XOR mask of the output map's shared-memory swizzle on the 16-byte chunk index (7 / 3 / 1 / 0 = 128B / 64B / 32B / none)
  int epi_tma;
  uint32_t epi_off, epi_swz;
  // resident-weight mode of conv_fprop_xslab1_kernel: the whole Toeplitz matrix (nt x kd*kh*kx) stays in shared memory for the
  // life of the CTA at byte offset wres_off; the ring then carries slab boxes only
  int wres;
  uint32_t wres_off;
  // variable-slot ring of conv_fprop_xslab1_kernel (one 64-box + one 32-box per dy): slots alternate big (128-byte rows) and
  // small (64-byte rows, packed) instead of all being big -- 4 stages in the space of 3
  int varslot;
};

constexpr int kSlabAWarps = 1, kSlabBWarps = 3;

__device__ __forceinline__ void xslab_decode(const XslabParams& p, int tile, int& n, int& z0, int& y0, int& g) {
  int t = tile;
  g = t % p.groups_x; t /= p.groups_x;
  y0 = (t % p.tiles_h) * 16; t /= p.tiles_h;
  z0 = (t % p.tiles_d) * 8 * p.rt; t /= p.tiles_d;
  n = t;
}

// Per-thread partial channel sums of the fused statistics (CB = Cout / 16 column blocks held in registers); folded with warp
// shuffles and added to `stats` with one fp64 atomic per (warp, channel, quantity) whenever the sample changes.
template <int CB>
struct EpiStats {
  static constexpr int NC = CB > 0 ? CB * 16 : 1;
  float s1[NC], s2[NC];
  int cur_n;
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < NC; ++i) s1[i] = s2[i] = 0.f;
    cur_n = -1;
  }
  __device__ __forceinline__ void flush(double* stats, int cout) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const float a = warp_sum(s1[i]), b = warp_sum(s2[i]);
      if (lane == 0) {
        atomicAdd(stats + ((int64_t)cur_n * cout + i) * 2, (double)a);
        atomicAdd(stats + ((int64_t)cur_n * cout + i) * 2 + 1, (double)b);
      }
      s1[i] = s2[i] = 0.f;
    }
  }
  __device__ __forceinline__ void sample(double* stats, int cout, int n) {
    if (n != cur_n) {
      if (cur_n >= 0) flush(stats, cout);
      cur_n = n;
    }
  }
  template <typename T>
  __device__ __forceinline__ void add16(int co, const Pack<T, 8>& w0, const Pack<T, 8>& w1) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float v = to_f<T>(i < 8 ? w0.v[i] : w1.v[i - 8]);
      s1[co + i] += v;
      s2[co + i] = fmaf(v, v, s2[co + i]);
    }
  }
};

// Direct-store epilogue of both x-slab kernels (warps 0-3, thread = one GEMM row = 4 x-voxels of one (y, z) line): TMEM ->
// +bias (-> +old value) -> T -> two 16-byte stores per 16 channels.  CB = 0: any Cout, no statistics.  CB = Cout / 16 in 1..3:
// the (j, channel-block) loop is unrolled so that the partial sums of EpiStats stay in registers.  Used when the output does
// not meet the alignment rules of the TMA-store epilogue below (or with B200_EPI_TMA=0).
template <typename T, int CB>
__device__ __forceinline__ void xslab_epilogue(const XslabParams& p, const float* __restrict__ bias, T* __restrict__ y, uint32_t tmem,
                                               uint32_t bar_tfull, uint32_t bar_tempty) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3;
  const int row = q * 32 + lane;
  const int ly = row & 15, lzr = row >> 4;
  EpiStats<CB> es;
  es.init();
  const bool stats = CB > 0 && p.stats != nullptr;
  int it = 0;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
    const int buf = p.nbuf == 2 ? (it & 1) : 0;
    const uint32_t par = p.nbuf == 2 ? ((it >> 1) & 1) : (it & 1);
    int n, z0, y0, g;
    xslab_decode(p, tile, n, z0, y0, g);
    if (stats) es.sample(p.stats, p.cout, n);
    {
      const long long c0 = p.dbg ? clock64() : 0;
      mbar_wait(bar_tfull + 8 * buf, par);
      if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 16 + 4] += clock64() - c0;
    }
    tc_fence_after();
    for (int r = 0; r < ((p.ablate & 4) ? 0 : p.rt); ++r) {
      const int gz = z0 + r * 8 + lzr, gy = y0 + ly;
      const bool valid = gz < p.d && gy < p.h;
      T* ybase = y + (int64_t)n * p.ysn + (int64_t)gz * p.ysd + (int64_t)gy * p.ysh + (int64_t)(4 * g) * p.ysw;
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * p.rt + r) * p.nt);
      auto chunk = [&](int c0, int j, int co) {
        uint32_t rr[16];
        tmem_ld16(taddr + c0, rr);
        tmem_ld_wait();
        if (valid && 4 * g + j < p.w) {
          T* yrow = ybase + (int64_t)j * p.ysw + co;
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(rr[i]) + (bias ? __ldg(bias + co + i) : 0.f);
          if (p.accumulate) {
            Pack<T, 8> o0 = *reinterpret_cast<const Pack<T, 8>*>(yrow);
            Pack<T, 8> o1 = *reinterpret_cast<const Pack<T, 8>*>(yrow + 8);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              f[i] += to_f<T>(o0.v[i]);
              f[8 + i] += to_f<T>(o1.v[i]);
            }
          }
          Pack<T, 8> w0, w1;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            w0.v[i] = from_f<T>(f[i]);
            w1.v[i] = from_f<T>(f[8 + i]);
          }
          *reinterpret_cast<Pack<T, 8>*>(yrow) = w0;
          *reinterpret_cast<Pack<T, 8>*>(yrow + 8) = w1;
          if (CB > 0 && stats) es.template add16<T>(CB > 0 ? co : 0, w0, w1);
        }
      };
      if (CB > 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int cb = 0; cb < (CB > 0 ? CB : 1); ++cb) chunk((j * CB + cb) * 16, j, cb * 16);
      } else {
        int j = 0, co = 0;
        for (int c0 = 0; c0 < p.nt; c0 += 16) {
          chunk(c0, j, co);
          co += 16;
          if (co == p.cout) { co = 0; ++j; }
        }
      }
    }
    tc_fence_before();
    mbar_arrive(bar_tempty + 8 * buf);
  }
  if (stats && es.cur_n >= 0) es.flush(p.stats, p.cout);
}

// TMA-store epilogue of the x-slab kernels.  The direct epilogue above has every thread store 16-byte pieces of its own
// (y, z) line: 32 lines per warp instruction, i.e. 32 LSU wavefronts and 32 half-written sectors per store, and the
// accumulate variant waits on a dependent global load per piece (B200: 16 -> 16 @128^3 0.280 ms plain, 0.435 ms
// accumulating; 16 -> 48: 0.90 / 1.59 ms).  Here a (row tile, 16-channel block) box -- 128 lines x 4 voxels x 16 channels =
// 16 KB -- is staged in shared memory in the layout of a rank-5 output map (C, W, H, D, N), box (16, 4, 16, 8, 1), and leaves
// with ONE bulk tensor store; `accumulate` becomes the element-wise add of the TMA unit (cp.reduce.async.bulk.tensor .add:
// the sum is formed in L2 on the rounded value, no read-back through the SM).  Two staging buffers alternate; TMEM is
// released as soon as the last column block of a tile sits in registers.  Same shapes: 0.248 / 0.268 ms and 0.73 / 0.78 ms.
template <typename T, int CB>
__device__ __forceinline__ void xslab_epilogue_tma(const XslabParams& p, const CUtensorMap* tmy, const float* __restrict__ bias,
                                                   uint32_t tmem, uint32_t bar_tfull, uint32_t bar_tempty, uint32_t stage) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3;
  const int row = q * 32 + lane;
  const int ly = row & 15, lzr = row >> 4;
  const int ncb = CB > 0 ? CB : p.cout / 16;
  EpiStats<CB> es;
  es.init();
  const bool stats = CB > 0 && p.stats != nullptr;
  const uint32_t swz = (uint32_t)row & p.epi_swz;
  const uint32_t srow = stage + (uint32_t)row * 128u;
  const bool issuer = threadIdx.x == 0;
  uint32_t nbox = 0;
  int it = 0;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
    const int buf = p.nbuf == 2 ? (it & 1) : 0;
    const uint32_t par = p.nbuf == 2 ? ((it >> 1) & 1) : (it & 1);
    int n, z0, y0, g;
    xslab_decode(p, tile, n, z0, y0, g);
    if (stats) es.sample(p.stats, p.cout, n);
    mbar_wait(bar_tfull + 8 * buf, par);
    tc_fence_after();
    for (int r = 0; r < p.rt; ++r) {
      const int gz = z0 + r * 8 + lzr, gy = y0 + ly;
      const bool valid = gz < p.d && gy < p.h;
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * p.rt + r) * p.nt);
      auto box = [&](int cb) {
        const int co = cb * 16;
        const uint32_t sbuf = srow + (nbox & 1u) * 16384u;
#pragma unroll
        for (int jp = 0; jp < 2; ++jp) {
          uint32_t rr[2][16];
          tmem_ld16(taddr + (uint32_t)(((2 * jp) * ncb + cb) * 16), rr[0]);
          tmem_ld16(taddr + (uint32_t)(((2 * jp + 1) * ncb + cb) * 16), rr[1]);
          tmem_ld_wait();
          if (jp == 0) {
            // staging buffer (nbox & 1) was last read by the store issued two boxes ago
            if (issuer) bulk_wait_group_read<1>();
            asm volatile("bar.sync 1, 128;" ::: "memory");
          } else if (r == p.rt - 1 && cb == ncb - 1) {
            tc_fence_before();
            mbar_arrive(bar_tempty + 8 * buf);     // the whole accumulator of this tile has been read
          }
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int j = 2 * jp + jj;
            Pack<T, 8> w0, w1;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              w0.v[i] = from_f<T>(__uint_as_float(rr[jj][i]) + (bias ? __ldg(bias + co + i) : 0.f));
              w1.v[i] = from_f<T>(__uint_as_float(rr[jj][8 + i]) + (bias ? __ldg(bias + co + 8 + i) : 0.f));
            }
            st_shared_v4(sbuf + (((uint32_t)(2 * j)) ^ swz) * 16u, *reinterpret_cast<const uint4*>(&w0));
            st_shared_v4(sbuf + (((uint32_t)(2 * j + 1)) ^ swz) * 16u, *reinterpret_cast<const uint4*>(&w1));
            if (CB > 0 && stats && valid && 4 * g + j < p.w) es.template add16<T>(CB > 0 ? co : 0, w0, w1);
          }
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (issuer) {
          if (p.accumulate) tma_reduce_add_5d(tmy, stage + (nbox & 1u) * 16384u, co, 4 * g, y0, z0 + r * 8, n);
          else tma_store_5d(tmy, stage + (nbox & 1u) * 16384u, co, 4 * g, y0, z0 + r * 8, n);
          bulk_commit_group();
        }
        ++nbox;
      };
      if (CB > 0) {
#pragma unroll
        for (int cb = 0; cb < (CB > 0 ? CB : 1); ++cb) box(cb);
      } else {
        for (int cb = 0; cb < ncb; ++cb) box(cb);
      }
    }
  }
  if (stats && es.cur_n >= 0) es.flush(p.stats, p.cout);
  if (issuer) bulk_wait_group_read<0>();
}

template <typename T, int CB>
__global__ void __launch_bounds__(32 * (6 + kSlabAWarps + kSlabBWarps), 1)
conv_fprop_xslab_kernel(const __grid_constant__ CUtensorMap tmx64, const __grid_constant__ CUtensorMap tmx32,
                        const __grid_constant__ CUtensorMap tmw64, const __grid_constant__ CUtensorMap tmw32,
                        const __grid_constant__ CUtensorMap tmy, const float* __restrict__ bias, T* __restrict__ y,
                        const XslabParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_bar[4 * kMaxStages + 4];
  __shared__ uint32_t s_tmem;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_afull = smem_u32(&s_bar[0]);
  const uint32_t bar_aempty = smem_u32(&s_bar[kMaxStages]);
  const uint32_t bar_bfull = smem_u32(&s_bar[2 * kMaxStages]);
  const uint32_t bar_bempty = smem_u32(&s_bar[3 * kMaxStages]);
  const uint32_t bar_tfull = smem_u32(&s_bar[4 * kMaxStages]);
  const uint32_t bar_tempty = smem_u32(&s_bar[4 * kMaxStages + 2]);

  if (threadIdx.x == 0) {
    // every row tile has its own MMA-issuing warp: slots are released and accumulators published by p.rt commits
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(bar_afull + 8 * s, 1); mbar_init(bar_aempty + 8 * s, p.rt); }
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(bar_bfull + 8 * s, 1); mbar_init(bar_bempty + 8 * s, p.rt); }
    for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, p.rt); mbar_init(bar_tempty + 8 * b, 128); }
    fence_barrier_init();
    tma_prefetch_desc(&tmx64); tma_prefetch_desc(&tmx32); tma_prefetch_desc(&tmw64); tma_prefetch_desc(&tmw32);
    if (p.epi_tma) tma_prefetch_desc(&tmy);
  }
  if (warp == 4) tmem_alloc(smem_u32(&s_tmem), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const int pd = p.kd / 2, ph_pad = p.kh / 2;
  const int nboxes = p.boxes64 + p.has32;
  const uint32_t slab_rows = (uint32_t)p.zl * 16u;

  auto decode = [&](int tile, int& n, int& z0, int& y0, int& g) {
    int t = tile;
    g = t % p.groups_x; t /= p.groups_x;
    y0 = (t % p.tiles_h) * 16; t /= p.tiles_h;
    z0 = (t % p.tiles_d) * 8 * p.rt; t /= p.tiles_d;
    n = t;
  };

  if (warp >= 6) {
    // =================================================================== TMA producers
    // One thread sustains only one bulk-tensor copy per ~450 cycles whatever the box size (tools/pipe_rates.py), so the
    // slab ring is fed by kSlabAWarps warps and the weight ring by kSlabBWarps warps, each taking every k-th item.
    if (!(p.ablate & 1) && elect_one()) {
      const bool a_role = warp < 6 + kSlabAWarps;
      const int rank = a_role ? warp - 6 : warp - 6 - kSlabAWarps;
      const int nrole = a_role ? kSlabAWarps : kSlabBWarps;
      int turn = 0, slot = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int n, z0, y0, g;
        decode(tile, n, z0, y0, g);
        const int e0 = (4 * g - p.kw / 2) * p.cin;
        for (int dy = 0; dy < p.kh; ++dy)
          for (int b = 0; b < nboxes; ++b) {
            const bool wide = b < p.boxes64;
            const uint32_t wbytes = wide ? 128u : 64u;
            if (a_role) {
              if (turn == rank) {
                mbar_wait(bar_aempty + 8 * slot, ph ^ 1);
                const uint32_t fa = bar_afull + 8 * slot;
                mbar_expect_tx(fa, slab_rows * wbytes);
                asm volatile(
                    "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                    ::"r"(smem0 + slot * p.a_bytes), "l"((uint64_t)(wide ? &tmx64 : &tmx32)), "r"(fa), "r"(e0 + b * 64),
                      "r"(y0 + dy - ph_pad), "r"(z0 - pd), "r"(n)
                    : "memory");
              }
              if (++turn == nrole) turn = 0;
              if (++slot == p.a_stages) { slot = 0; ph ^= 1; }
            } else {
              for (int dz = 0; dz < p.kd; ++dz) {
                if (turn == rank) {
                  mbar_wait(bar_bempty + 8 * slot, ph ^ 1);
                  const uint32_t fb = bar_bfull + 8 * slot;
                  mbar_expect_tx(fb, (uint32_t)p.nt * wbytes);
                  tma_load_2d(smem0 + p.b_off + slot * p.b_bytes, wide ? &tmw64 : &tmw32, fb, (dz * p.kh + dy) * p.kx + b * 64, 0);
                }
                if (++turn == nrole) turn = 0;
                if (++slot == p.b_stages) { slot = 0; ph ^= 1; }
              }
            }
          }
      }
    }
  } else if (warp >= 4) {
    const int r = warp - 4;                       // row tile issued by this warp
    if (r < p.rt && elect_one()) {
      // ================================================================= MMA issuers (one warp per row tile)
      // The issuing thread is the critical path of the N = 64 layers (ncu: ~8 cycles per dependent scalar instruction,
      // tensor pipe 30 % busy), so: loop constants pinned in registers, descriptors = template + slot offset, the
      // barrier try_wait of the NEXT weight stage is issued before the current stage's MMAs, and the two row tiles
      // that share every operand stage are issued by two warps.
      const int kh = in_reg(p.kh), kd = in_reg(p.kd), a_stages = in_reg(p.a_stages), b_stages = in_reg(p.b_stages);
      const int boxes64 = in_reg(p.boxes64), nb = in_reg(nboxes), nbuf = in_reg(p.nbuf), num_tiles = in_reg(p.num_tiles);
      const uint32_t nt = in_reg((uint32_t)p.nt), idesc = in_reg(p.idesc);
      const uint32_t a16 = in_reg(p.a_bytes >> 4), b16 = in_reg(p.b_bytes >> 4);
      const uint64_t tmpl128 = make_smem_desc(0, 16, 1024, kSwizzle128), tmpl64 = make_smem_desc(0, 16, 512, kSwizzle64);
      const uint32_t a0 = in_reg(smem0 >> 4), b0 = in_reg((smem0 + p.b_off) >> 4);
      const uint32_t rt_n = in_reg((uint32_t)p.rt * nt), stride = in_reg((int)gridDim.x);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int it = 0;
      const bool no_tma = p.ablate & 1, no_mma = p.ablate & 2;
      bool b_ready = no_tma || mbar_try_wait(bar_bfull, 0);
      long long c_tempty = 0, c_afull = 0, c_bfull = 0;
      const bool dbg = p.dbg != nullptr;
      const long long c_start = clock64();
      for (int tile = blockIdx.x; tile < num_tiles; tile += stride, ++it) {
        const int buf = nbuf == 2 ? (it & 1) : 0;
        const uint32_t par = nbuf == 2 ? ((it >> 1) & 1) : (it & 1);
        {
          const long long c0 = dbg ? clock64() : 0;
          mbar_wait(bar_tempty + 8 * buf, par ^ 1);
          if (dbg) c_tempty += clock64() - c0;
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem + (uint32_t)buf * rt_n + (uint32_t)r * nt;
        uint32_t acc = 0;
        for (int dy = 0; dy < kh; ++dy)
          for (int b = 0; b < nb; ++b) {
            const bool wide = b < boxes64;
            // one 16-row line of the slab in 16-byte units (16 x 128 B or 16 x 64 B); row tile r starts 8 lines in
            const uint32_t line = wide ? 128u : 64u;
            const uint64_t tmpl = wide ? tmpl128 : tmpl64;
            if (!no_tma) {
              const long long c0 = dbg ? clock64() : 0;
              mbar_wait(bar_afull + 8 * as, aph);
              if (dbg) c_afull += clock64() - c0;
            }
            uint64_t ad = tmpl + (uint64_t)(a0 + (uint32_t)as * a16 + (uint32_t)r * 8u * line);
            for (int dz = 0; dz < kd; ++dz, ad += line) {
              if (!b_ready) {
                const long long c0 = dbg ? clock64() : 0;
                mbar_wait(bar_bfull + 8 * bs, bph);
                if (dbg) c_bfull += clock64() - c0;
              }
              tc_fence_after();
              const uint64_t bd = tmpl + (uint64_t)(b0 + (uint32_t)bs * b16);
              const uint32_t cur = bar_bempty + 8 * bs;
              if (++bs == b_stages) { bs = 0; bph ^= 1; }
              b_ready = no_tma || mbar_try_wait(bar_bfull + 8 * bs, bph);          // overlaps the issue below
              if (no_mma) {
              } else if (wide) umma_ksteps<4>(d_tmem, ad, bd, idesc, acc);
              else umma_ksteps<2>(d_tmem, ad, bd, idesc, acc);
              acc = 1;
              umma_commit(cur);
            }
            umma_commit(bar_aempty + 8 * as);
            if (++as == a_stages) { as = 0; aph ^= 1; }
          }
        umma_commit(bar_tfull + 8 * buf);
      }
      if (dbg) {
        long long* o = p.dbg + (blockIdx.x * 2 + r) * 8;
        o[0] = clock64() - c_start; o[1] = c_tempty; o[2] = c_afull; o[3] = c_bfull;
      }
    }
  } else {
    // =================================================================== epilogue
    if (p.epi_tma) xslab_epilogue_tma<T, CB>(p, &tmy, bias, tmem, bar_tfull, bar_tempty, smem0 + p.epi_off);
    else xslab_epilogue<T, CB>(p, bias, y, tmem, bar_tfull, bar_tempty);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, p.tmem_cols);
  }
}

// Unified-stage variant for the narrow layers (kd * N * 128 B of weights per (dy, box) fit next to the slab): one stage =
// slab box + the kd weight tiles that multiply it, one full/empty barrier pair per stage.  The issuing threads are the
// critical path of the N = 64 layers -- a stage hand-over costs ~250 cycles of dependent scalar instructions however
// little it moves (B200_ABLATE / B200_DBG measurements, profiles/README.md) -- so a CTA tile takes kh * boxes (6 for
// 16 -> 16) hand-overs here instead of kh * boxes * (kd + 1) (24) in conv_fprop_xslab_kernel.
template <typename T, int CB>
__global__ void __launch_bounds__(320, 1)
conv_fprop_xslab1_kernel(const __grid_constant__ CUtensorMap tmx64, const __grid_constant__ CUtensorMap tmx32,
                         const __grid_constant__ CUtensorMap tmw64, const __grid_constant__ CUtensorMap tmw32,
                         const __grid_constant__ CUtensorMap tmy, const float* __restrict__ bias, T* __restrict__ y,
                         const XslabParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_bar[4 * kMaxStages + 4];
  __shared__ uint32_t s_tmem;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_afull = smem_u32(&s_bar[0]);
  const uint32_t bar_aempty = smem_u32(&s_bar[kMaxStages]);
  const uint32_t bar_wres = smem_u32(&s_bar[2 * kMaxStages]);
  const uint32_t bar_tfull = smem_u32(&s_bar[4 * kMaxStages]);
  const uint32_t bar_tempty = smem_u32(&s_bar[4 * kMaxStages + 2]);
  const bool wres = p.wres != 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(bar_afull + 8 * s, wres ? 1 : 1 + p.kd); mbar_init(bar_aempty + 8 * s, p.rt); }
    if (wres) mbar_init(bar_wres, p.kd);
    for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, p.rt); mbar_init(bar_tempty + 8 * b, 128); }
    fence_barrier_init();
    tma_prefetch_desc(&tmx64); tma_prefetch_desc(&tmx32); tma_prefetch_desc(&tmw64); tma_prefetch_desc(&tmw32);
    if (p.epi_tma) tma_prefetch_desc(&tmy);
  }
  if (warp == 4) tmem_alloc(smem_u32(&s_tmem), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const int pd = p.kd / 2, ph_pad = p.kh / 2;
  const int nboxes = p.boxes64 + p.has32;
  const uint32_t slab_rows = (uint32_t)p.zl * 16u;
  // weight tiles follow the (max-size) slab box inside a stage; rowb = bytes per row of the widest box (128, or 64 when the
  // window is cut into 32-element boxes only)
  const uint32_t rowb = p.boxes64 > 0 ? 128u : 64u;
  const uint32_t w_off = slab_rows * rowb;
  const uint32_t w_tile = (uint32_t)p.nt * rowb;
  // variable slots: even slots hold a 64-box stage (slab + kd tiles, 128-byte rows), odd slots a 32-box stage packed at 64-byte
  // rows; slot s starts at (s / 2) * (big + small) + (s & 1) * big
  const bool varslot = p.varslot != 0;
  const uint32_t vs_big = slab_rows * 128u + (uint32_t)p.kd * (uint32_t)p.nt * 128u, vs_small = vs_big >> 1;
  // resident weights: tile (dy, b, dz) at wres_off + dy * wres_dy + (b < boxes64 ? b * kd * t64 : boxes64 * kd * t64) + dz * (t64 | t32)
  const uint32_t t64 = (uint32_t)p.nt * 128u, t32 = (uint32_t)p.nt * 64u;
  const uint32_t wres_dy = (uint32_t)p.kd * ((uint32_t)p.boxes64 * t64 + (uint32_t)p.has32 * t32);

  auto decode = [&](int tile, int& n, int& z0, int& y0, int& g) {
    int t = tile;
    g = t % p.groups_x; t /= p.groups_x;
    y0 = (t % p.tiles_h) * 16; t /= p.tiles_h;
    z0 = (t % p.tiles_d) * 8 * p.rt; t /= p.tiles_d;
    n = t;
  };

  if (warp >= 6) {
    // =================================================================== TMA producers: warp 6 = slab, warp 7 + dz = weight tile dz
    const int role = warp - 6;
    if (wres && role >= 1 && role <= p.kd && elect_one()) {
      // resident weights: warp 7 + dz fetches every (dy, box) tile of its dz once; all kd warps signal one barrier
      const int dz = role - 1;
      mbar_expect_tx(bar_wres, (uint32_t)p.kh * ((uint32_t)p.boxes64 * t64 + (uint32_t)p.has32 * t32));
      for (int dy = 0; dy < p.kh; ++dy)
        for (int b = 0; b < nboxes; ++b) {
          const bool wide = b < p.boxes64;
          const uint32_t dst = smem0 + p.wres_off + (uint32_t)dy * wres_dy +
                               (wide ? (uint32_t)b * p.kd * t64 : (uint32_t)p.boxes64 * p.kd * t64 + (uint32_t)(b - p.boxes64) * p.kd * t32) +
                               (uint32_t)dz * (wide ? t64 : t32);
          tma_load_2d(dst, wide ? &tmw64 : &tmw32, bar_wres, (dz * p.kh + dy) * p.kx + (wide ? b * 64 : p.boxes64 * 64 + (b - p.boxes64) * 32), 0);
        }
    } else if (role <= (wres ? 0 : p.kd) && elect_one()) {
      const int a_stages = p.a_stages;
      int slot = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int n, z0, y0, g;
        decode(tile, n, z0, y0, g);
        const int e0 = (4 * g - p.kw / 2 - p.xoff) * p.cin;
        for (int dy = 0; dy < p.kh; ++dy)
          for (int b = 0; b < nboxes; ++b) {
            const bool wide = b < p.boxes64;
            const uint32_t wbytes = wide ? 128u : 64u;
            const int koff = wide ? b * 64 : p.boxes64 * 64 + (b - p.boxes64) * 32;      // element offset of box b in the window
            mbar_wait(bar_aempty + 8 * slot, ph ^ 1);
            const uint32_t fa = bar_afull + 8 * slot;
            const uint32_t dst = smem0 + (varslot ? (uint32_t)(slot >> 1) * (vs_big + vs_small) + (uint32_t)(slot & 1) * vs_big
                                                  : (uint32_t)slot * p.a_bytes);
            const uint32_t w_off_b = (varslot && !wide) ? slab_rows * 64u : w_off;
            const uint32_t w_tile_b = (varslot && !wide) ? (uint32_t)p.nt * 64u : w_tile;
            if (role == 0) {
              mbar_expect_tx(fa, slab_rows * wbytes);
              asm volatile(
                  "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                  ::"r"(dst), "l"((uint64_t)(wide ? &tmx64 : &tmx32)), "r"(fa), "r"(e0 + koff), "r"(y0 + dy - ph_pad),
                    "r"(z0 - pd), "r"(n)
                  : "memory");
            } else {
              const int dz = role - 1;
              mbar_expect_tx(fa, (uint32_t)p.nt * wbytes);
              tma_load_2d(dst + w_off_b + (uint32_t)dz * w_tile_b, wide ? &tmw64 : &tmw32, fa, (dz * p.kh + dy) * p.kx + koff, 0);
            }
            if (++slot == a_stages) { slot = 0; ph ^= 1; }
          }
      }
    }
  } else if (warp >= 4) {
    const int r = warp - 4;                       // row tile issued by this warp
    if (r < p.rt && elect_one()) {
      // ================================================================= MMA issuers (one warp per row tile)
      const int kh = in_reg(p.kh), kd = in_reg(p.kd), a_stages = in_reg(p.a_stages);
      const int boxes64 = in_reg(p.boxes64), nb = in_reg(nboxes), nbuf = in_reg(p.nbuf), num_tiles = in_reg(p.num_tiles);
      const uint32_t nt = in_reg((uint32_t)p.nt), idesc = in_reg(p.idesc);
      const uint32_t a16 = in_reg(p.a_bytes >> 4), w16 = in_reg(w_off >> 4), wt16 = in_reg(w_tile >> 4);
      const uint64_t tmpl128 = make_smem_desc(0, 16, 1024, kSwizzle128), tmpl64 = make_smem_desc(0, 16, 512, kSwizzle64);
      const uint32_t a0 = in_reg(smem0 >> 4);
      const uint32_t rt_n = in_reg((uint32_t)p.rt * nt), stride = in_reg((int)gridDim.x);
      const uint32_t wres16 = in_reg((smem0 + p.wres_off) >> 4), wres_dy16 = in_reg(wres_dy >> 4), t64_16 = in_reg(t64 >> 4),
                     t32_16 = in_reg(t32 >> 4);
      const uint32_t vs_big16 = in_reg(vs_big >> 4), vs_small16 = in_reg(vs_small >> 4);
      const uint32_t w16s = in_reg((slab_rows * 64u) >> 4), wt16s = in_reg(((uint32_t)p.nt * 64u) >> 4);
      int as = 0;
      uint32_t aph = 0;
      int it = 0;
      if (wres) {
        mbar_wait(bar_wres, 0);
        tc_fence_after();
      }
      bool ready = mbar_try_wait(bar_afull, 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += stride, ++it) {
        const int buf = nbuf == 2 ? (it & 1) : 0;
        const uint32_t par = nbuf == 2 ? ((it >> 1) & 1) : (it & 1);
        mbar_wait(bar_tempty + 8 * buf, par ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem + (uint32_t)buf * rt_n + (uint32_t)r * nt;
        uint32_t acc = 0;
        for (int dy = 0; dy < kh; ++dy)
          for (int b = 0; b < nb; ++b) {
            const bool wide = b < boxes64;
            const uint32_t line = wide ? 128u : 64u;         // one 16-row slab line in 16-byte units
            const uint64_t tmpl = wide ? tmpl128 : tmpl64;
            if (!ready) mbar_wait(bar_afull + 8 * as, aph);
            tc_fence_after();
            const uint32_t s16 = varslot ? a0 + (uint32_t)(as >> 1) * (vs_big16 + vs_small16) + (uint32_t)(as & 1) * vs_big16
                                         : a0 + (uint32_t)as * a16;
            const uint32_t t_lo = (uint32_t)tmpl, t_hi = (uint32_t)(tmpl >> 32);
            uint32_t ad = t_lo + s16 + (uint32_t)r * 8u * line;
            uint32_t bd = t_lo + s16 + ((varslot && !wide) ? w16s : w16);
            uint32_t bstep = (varslot && !wide) ? wt16s : wt16;
            if (wres) {
              bd = t_lo + wres16 + (uint32_t)dy * wres_dy16 +
                   (wide ? (uint32_t)b * (uint32_t)kd * t64_16 : (uint32_t)boxes64 * (uint32_t)kd * t64_16 + (uint32_t)(b - boxes64) * (uint32_t)kd * t32_16);
              bstep = wide ? t64_16 : t32_16;
            }
            const uint32_t cur = bar_aempty + 8 * as;
            if (++as == a_stages) { as = 0; aph ^= 1; }
            ready = mbar_try_wait(bar_afull + 8 * as, aph);              // latency overlaps the issue below
            for (int dz = 0; dz < kd; ++dz, ad += line, bd += bstep) {
              if (wide) umma_ksteps_split<4>(d_tmem, ad, t_hi, bd, t_hi, 2u, 2u, idesc, acc);
              else umma_ksteps_split<2>(d_tmem, ad, t_hi, bd, t_hi, 2u, 2u, idesc, acc);
              acc = 1;
            }
            umma_commit(cur);
          }
        umma_commit(bar_tfull + 8 * buf);
      }
    }
  } else {
    // =================================================================== epilogue
    if (p.epi_tma) xslab_epilogue_tma<T, CB>(p, &tmy, bias, tmem, bar_tfull, bar_tempty, smem0 + p.epi_off);
    else xslab_epilogue<T, CB>(p, bias, y, tmem, bar_tfull, bar_tempty);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, p.tmem_cols);
  }
}

// Toeplitz packing: out[(j, co)][(dz, dy, xi, ci)] from w (Cout, Cin, kd, kh, 3) fp32; flip_transpose builds the dgrad
// operand (roles of Cin/Cout swapped, taps mirrored) directly.
// Window of the x-folded row for (Cin, kw): it starts `xoff` voxels before the first tap (4g - kw/2 - xoff) and holds kxp
// elements per (dz, dy).  Cin a multiple of 16: xoff = 0, kxp = (3 + kw) * Cin.  Image-fed layers (Cin = 2, 4, 8) get the
// smallest xoff that makes the window start 16-byte aligned and a length padded to whole 32-element TMA boxes
// (Cin = 2, kw = 3: xoff = 3, 9 voxels -> 32 elements, i.e. two K steps per (dz, dy) instead of the six of a 16-padded input).
__host__ __device__ inline bool xfold_geom(int cin, int kw, int* xoff, int* kxp) {
  const int pw = kw / 2;
  if (cin % 16 == 0) {
    *xoff = 0;
    *kxp = (3 + kw) * cin;
    return *kxp % 32 == 0;
  }
  if (cin != 2 && cin != 4 && cin != 8) return false;
  int xo = 0;
  while (((pw + xo) * cin * 2) % 16 != 0) ++xo;
  *xoff = xo;
  *kxp = ((xo + 4 + 2 * pw) * cin + 31) / 32 * 32;
  return true;
}

template <typename T>
__global__ void pack_weight_xfold_kernel(const float* __restrict__ w, T* __restrict__ out, int cout, int cin, int kd, int kh,
                                         int kw, int flip) {
  // logical conv: co_l in [0, CO), ci_l in [0, CI) with CO = flip ? cin : cout, CI = flip ? cout : cin
  const int CO = flip ? cin : cout, CI = flip ? cout : cin;
  int xoff = 0, kxp = 0;
  xfold_geom(CI, kw, &xoff, &kxp);
  const int win = kxp / CI;                              // voxels per (dz, dy) window, padding included
  const int64_t ktot = (int64_t)kd * kh * kxp;
  const int64_t total = (int64_t)4 * CO * ktot;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int ci = (int)(t % CI); t /= CI;
    const int xi = (int)(t % win); t /= win;
    const int dy = (int)(t % kh); t /= kh;
    const int dz = (int)(t % kd); t /= kd;
    const int co = (int)(t % CO);
    const int j = (int)(t / CO);
    const int kx = xi - xoff - j;
    float v = 0.f;
    if (kx >= 0 && kx < kw) {
      if (!flip) v = w[((((int64_t)co * cin + ci) * kd + dz) * kh + dy) * kw + kx];
      else v = w[((((int64_t)ci * cin + co) * kd + (kd - 1 - dz)) * kh + (kh - 1 - dy)) * kw + (kw - 1 - kx)];
    }
    out[i] = from_f<T>(v);
  }
}

// ------------------------------------------------------------------------------------- fprop kernel, cp.async-fed
// Same GEMM and epilogue as conv_fprop_umma_kernel, but the operands are brought in by four loader warps with
// 16-byte cp.async (zero-fill for padding) straight into the swizzled K-major layout the UMMA descriptors expect.
// Measured reason (profiles/README.md): a TMA box is fetched one innermost row at a time (~6 cycles per row,
// independent of the row size), so the 32..64-byte channel rows of the C <= 32 layers ran at ~6 B/clk/SM; the LSU path
// moves the same bytes with a handful of instructions per thread and lets one pipeline stage carry all kw taps.
// Warp roles (288 threads): warp 0 = MMA issuer + TMEM owner, warps 1-4 = epilogue, warps 5-8 = loaders.
struct FpropParams2 {
  int n, d, h, w, cin, cout;
  int kd, kh, kw;
  int bd, bh, bw;
  int tiles_d, tiles_h, tiles_w, tiles_n;
  int num_tiles;
  int ck, chunks;
  int nt;
  int tg;                               // taps per pipeline stage (kw or 1)
  int stages;
  uint32_t a_bytes, b_bytes, stage_bytes;   // per tap / per tap / per stage (tg taps)
  uint32_t layout, sbo;
  uint32_t idesc;
  uint32_t tmem_cols;
  int64_t xsw, xsh, xsd, xsn;           // input element strides
  int64_t ysw, ysh, ysd, ysn;           // output element strides
  int accumulate;
};

__device__ __forceinline__ uint32_t swz_chunk(uint32_t row, uint32_t j, int ck) {
  // 16-byte chunk j of row `row` inside a K-major tile with ck*2-byte rows (Swizzle<1|2|3,4,3> of the byte offset)
  return ck == 64 ? (j ^ (row & 7u)) : (ck == 32 ? (j ^ ((row >> 1) & 3u)) : (j ^ ((row >> 2) & 1u)));
}

template <typename T>
__global__ void __launch_bounds__(288, 1)
conv_fprop_umma2_kernel(const T* __restrict__ x, const T* __restrict__ wp, const float* __restrict__ bias, T* __restrict__ y,
                        const FpropParams2 p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_bar[2 * kMaxStages + 4];
  __shared__ uint32_t s_tmem;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = smem_u32(&s_bar[0]);
  const uint32_t bar_empty = smem_u32(&s_bar[kMaxStages]);
  const uint32_t bar_tfull = smem_u32(&s_bar[2 * kMaxStages]);
  const uint32_t bar_tempty = smem_u32(&s_bar[2 * kMaxStages + 2]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 128);     // one asynchronous arrival per loader thread
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, 128);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  const int taps = p.kd * p.kh * p.kw;
  const int groups = taps / p.tg;                 // tap groups (tg divides kw)
  const int num_kb = groups * p.chunks;
  const int pd = p.kd / 2, ph = p.kh / 2, pw = p.kw / 2;

  auto decode = [&](int tile, int& n, int& z0, int& y0, int& x0, int& n0) {
    int t = tile;
    n0 = (t % p.tiles_n) * p.nt; t /= p.tiles_n;
    x0 = (t % p.tiles_w) * p.bw; t /= p.tiles_w;
    y0 = (t % p.tiles_h) * p.bh; t /= p.tiles_h;
    z0 = (t % p.tiles_d) * p.bd; t /= p.tiles_d;
    n = t;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ================================================================= MMA issuer
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(bar_tempty + 8 * buf, ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem + buf * p.nt;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          fence_proxy_async();
          tc_fence_after();
          const uint32_t s_base = smem0 + stage * p.stage_bytes;
          for (int t = 0; t < p.tg; ++t) {
            const uint32_t a_base = s_base + t * (p.a_bytes + p.b_bytes);
            const uint32_t b_base = a_base + p.a_bytes;
            for (int k = 0; k < p.ck / 16; ++k) {
              const uint64_t ad = make_smem_desc(a_base + k * 32, 16, p.sbo, p.layout);
              const uint64_t bd = make_smem_desc(b_base + k * 32, 16, p.sbo, p.layout);
              umma_f16(d_tmem, ad, bd, p.idesc, (kb | t | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(bar_empty + 8 * stage);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(bar_tfull + 8 * buf);
      }
    }
  } else if (warp <= 4) {
    // =================================================================== epilogue (TMEM lane quarter = warp % 4)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int lx = row % p.bw, ly = (row / p.bw) % p.bh, lz = row / (p.bw * p.bh);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      int n, z0, y0, x0, n0;
      decode(tile, n, z0, y0, x0, n0);
      mbar_wait(bar_tfull + 8 * buf, (it >> 1) & 1);
      tc_fence_after();
      const int gz = z0 + lz, gy = y0 + ly, gx = x0 + lx;
      const bool valid = gz < p.d && gy < p.h && gx < p.w;
      T* yrow = y + (int64_t)n * p.ysn + (int64_t)gz * p.ysd + (int64_t)gy * p.ysh + (int64_t)gx * p.ysw + n0;
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * p.nt);
      for (int j0 = 0; j0 < p.nt; j0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + j0, r);
        tmem_ld_wait();
        if (valid && n0 + j0 < p.cout) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(r[j]) + (bias ? __ldg(bias + n0 + j0 + j) : 0.f);
          if (p.accumulate) {
            Pack<T, 8> o0 = *reinterpret_cast<const Pack<T, 8>*>(yrow + j0);
            Pack<T, 8> o1 = *reinterpret_cast<const Pack<T, 8>*>(yrow + j0 + 8);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              f[j] += to_f<T>(o0.v[j]);
              f[8 + j] += to_f<T>(o1.v[j]);
            }
          }
          Pack<T, 8> w0, w1;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            w0.v[j] = from_f<T>(f[j]);
            w1.v[j] = from_f<T>(f[8 + j]);
          }
          *reinterpret_cast<Pack<T, 8>*>(yrow + j0) = w0;
          *reinterpret_cast<Pack<T, 8>*>(yrow + j0 + 8) = w1;
        }
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * buf);
    }
  } else {
    // =================================================================== loaders (128 threads, one A row each)
    const int r = threadIdx.x - 160;
    const int lx = r % p.bw, ly = (r / p.bw) % p.bh, lz = r / (p.bw * p.bh);
    const int cpr = p.ck / 8;                       // 16-byte chunks per row (2, 4 or 8)
    const int cpr_shift = p.ck == 16 ? 1 : (p.ck == 32 ? 2 : 3);
    const int b_chunks = p.nt * cpr;                // 16-byte chunks of one weight tile
    const int64_t wrow = (int64_t)taps * p.cin;     // packed weight row length (elements)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int n, z0, y0, x0, n0;
      decode(tile, n, z0, y0, x0, n0);
      const int gz = z0 + lz, gy = y0 + ly, gx = x0 + lx;
      const T* xrow = x + (int64_t)n * p.xsn + (int64_t)gz * p.xsd + (int64_t)gy * p.xsh + (int64_t)gx * p.xsw;
      const uint32_t row_off = (uint32_t)r * (uint32_t)(p.ck * 2);
      // stage = one (dz, dy) pair x `tg` consecutive dx taps x one channel chunk
      for (int dz = 0; dz < p.kd; ++dz) {
        const int iz = gz + dz - pd;
        const bool okz = iz >= 0 && iz < p.d;
        for (int dy = 0; dy < p.kh; ++dy) {
          const int iy = gy + dy - ph;
          const bool okzy = okz && iy >= 0 && iy < p.h;
          const T* line = xrow + (int64_t)(dz - pd) * p.xsd + (int64_t)(dy - ph) * p.xsh;
          const int tap0 = (dz * p.kh + dy) * p.kw;
          for (int dx0 = 0; dx0 < p.kw; dx0 += p.tg)
            for (int ch = 0; ch < p.chunks; ++ch) {
              mbar_wait(bar_empty + 8 * stage, phase ^ 1);
              uint32_t a_base = smem0 + stage * p.stage_bytes;
              for (int t = 0; t < p.tg; ++t, a_base += p.a_bytes + p.b_bytes) {
                const int dx = dx0 + t;
                const int ix = gx + dx - pw;
                const bool ok = okzy && ix >= 0 && ix < p.w;
                const T* src = ok ? line + (int64_t)(dx - pw) * p.xsw + ch * p.ck : x;
                const uint32_t arow = a_base + row_off;
                const uint32_t nb = ok ? 16u : 0u;
                if (p.ck == 16) {
                  const uint32_t sw = ((uint32_t)r >> 2) & 1u;
                  cp_async16(arow + (sw << 4), src, nb);
                  cp_async16(arow + ((sw ^ 1u) << 4), src + 8, nb);
                } else {
                  for (int j = 0; j < cpr; ++j) cp_async16(arow + (swz_chunk((uint32_t)r, (uint32_t)j, p.ck) << 4), src + j * 8, nb);
                }
                // weight tile [nt][ck] of this tap / chunk: 16-byte chunk c -> (row c / cpr, chunk c % cpr)
                const uint32_t b_base = a_base + p.a_bytes;
                const T* wt = wp + (int64_t)(tap0 + dx) * p.cin + ch * p.ck;
                for (int c = r; c < b_chunks; c += 128) {
                  const int brow = c >> cpr_shift, j = c & (cpr - 1);
                  const bool bok = n0 + brow < p.cout;
                  cp_async16(b_base + (uint32_t)brow * (uint32_t)(p.ck * 2) + (swz_chunk((uint32_t)brow, (uint32_t)j, p.ck) << 4),
                             bok ? wt + (int64_t)(n0 + brow) * wrow + j * 8 : wp, bok ? 16u : 0u);
                }
              }
              cp_async_arrive_noinc(bar_full + 8 * stage);
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, p.tmem_cols);
  }
}

static void pick_tile(int d, int h, int w, int* bd, int* bh, int* bw) {
  static const int cand[][3] = {{2, 8, 8}, {4, 4, 8}, {1, 8, 16}, {1, 16, 8}, {2, 4, 16}, {4, 8, 4}, {8, 4, 4}, {1, 4, 32},
                                {2, 2, 32}, {1, 2, 64}, {1, 1, 128}, {2, 16, 4}, {1, 32, 4}, {4, 2, 16}, {8, 8, 2}, {16, 8, 1},
                                {1, 64, 2}, {1, 128, 1}, {128, 1, 1}, {2, 64, 1}, {8, 16, 1}, {4, 32, 1}, {32, 4, 1}, {64, 2, 1}};
  double best = 1e30;
  for (auto& c : cand) {
    double cover = (double)ceil_div(d, c[0]) * c[0] * ceil_div(h, c[1]) * c[1] * ceil_div(w, c[2]) * c[2];
    if (cover < best - 0.5) {
      best = cover;
      *bd = c[0]; *bh = c[1]; *bw = c[2];
    }
  }
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

static int loader_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("B200_CONV_LOADER");
    mode = (e && strcmp(e, "cpasync") == 0) ? 1 : 0;   // default: TMA (measured faster, profiles/README.md)
  }
  return mode;
}

struct Upscale { int sd = 0, sh = 0, sw = 0, cout = 0; };   // transposed-convolution scatter epilogue (off when sd == 0)
int conv_fprop_umma_v(const ActView& x, const void* w, const float* bias, const ActView& y, int kd, int kh, int kw,
                      int accumulate, cudaStream_t st, Upscale up = Upscale());
int conv_wgrad_umma_v(const ActView& x, const ActView& dy, float* dw, int kd, int kh, int kw, cudaStream_t st);
int conv_fprop_phases_v(const ActView* phase_views, int nphases, const void* w, const ActView& y, int accumulate, cudaStream_t st);
int conv_wgrad_phases_v(const ActView& x, const ActView* phase_views, int nphases, float* dw, cudaStream_t st);
bool conv_wgrad_xfold_ok(const ActView& x, const ActView& dy, int kd, int kh, int kw);
int conv_wgrad_xfold_v(const ActView& x, const ActView& dy, float* dw, int kd, int kh, int kw, cudaStream_t st);

}  // namespace sm100

bool conv_fprop_umma_supported(const b200_tensor* x, const b200_tensor* res, const b200_tensor* y, int kd, int kh, int kw) {
  if (x->dtype != B200_BF16 && x->dtype != B200_F16) return false;
  if (res != nullptr) return false;
  if (x->c % 16 != 0 || y->c % 16 != 0) return false;
  if (x->ld % 8 != 0 || y->ld % 8 != 0) return false;
  if (!sm100::aligned16(x->data) || !sm100::aligned16(y->data)) return false;
  if (kd * kh * kw > 125) return false;
  if ((int64_t)x->n * x->d * x->h * x->w < 128) return false;
  (void)kd; (void)kh; (void)kw;
  return sm100::encode_tiled_fn() != nullptr;
}

int conv_fprop_umma(const b200_tensor* x, const void* w, const float* bias, const b200_tensor* res, const b200_tensor* y,
                    int kd, int kh, int kw, int accumulate, cudaStream_t st) {
  (void)res;
  return sm100::conv_fprop_umma_v(sm100::view_of(x), w, bias, sm100::view_of(y), kd, kh, kw, accumulate, st);
}

namespace sm100 {
static int conv_fprop_umma2_v(const ActView& x, const void* w, const float* bias, const ActView& y, int kd, int kh, int kw,
                              int accumulate, cudaStream_t st) {
  FpropParams2 p{};
  p.n = x.n; p.d = x.d; p.h = x.h; p.w = x.w; p.cin = x.c; p.cout = y.c;
  p.kd = kd; p.kh = kh; p.kw = kw;
  pick_tile(x.d, x.h, x.w, &p.bd, &p.bh, &p.bw);
  p.ck = (x.c % 64 == 0) ? 64 : ((x.c % 32 == 0) ? 32 : 16);
  p.chunks = x.c / p.ck;
  p.nt = y.c <= 256 ? y.c : 128;
  p.tiles_d = (int)ceil_div(x.d, p.bd); p.tiles_h = (int)ceil_div(x.h, p.bh); p.tiles_w = (int)ceil_div(x.w, p.bw);
  p.tiles_n = (int)ceil_div(y.c, p.nt);
  p.num_tiles = x.n * p.tiles_d * p.tiles_h * p.tiles_w * p.tiles_n;
  p.a_bytes = 128u * p.ck * 2;
  p.b_bytes = (((uint32_t)p.nt * p.ck * 2) + 1023u) & ~1023u;       // keep every operand tile 1 KB aligned
  p.tg = (kw > 1 && (uint32_t)kw * (p.a_bytes + p.b_bytes) <= 24u * 1024u) ? kw : 1;
  p.stage_bytes = (uint32_t)p.tg * (p.a_bytes + p.b_bytes);
  int stages = (int)((200u * 1024u) / p.stage_bytes);
  if (stages > 8) stages = 8;
  B200_CHECK_ARG(stages >= 2, "conv_fprop(umma): tile does not fit in shared memory");
  p.stages = stages;
  p.layout = p.ck == 64 ? kSwizzle128 : (p.ck == 32 ? kSwizzle64 : kSwizzle32);
  p.sbo = 8u * p.ck * 2;
  p.idesc = make_idesc(x.dtype == B200_BF16, p.nt, 0, 0);
  uint32_t cols = 32;
  while (cols < 2u * p.nt) cols <<= 1;
  p.tmem_cols = cols;
  p.xsw = x.sw; p.xsh = x.sh; p.xsd = x.sd; p.xsn = x.sn;
  p.ysw = y.sw; p.ysh = y.sh; p.ysd = y.sd; p.ysn = y.sn;
  p.accumulate = accumulate;
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
  int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  if (x.dtype == B200_BF16) {
    auto kern = conv_fprop_umma2_kernel<__nv_bfloat16>;
    B200_CUDA(raise_dyn_smem_cap(kern));
    kern<<<grid, 288, smem, st>>>((const __nv_bfloat16*)x.data, (const __nv_bfloat16*)w, bias, (__nv_bfloat16*)y.data, p);
  } else {
    auto kern = conv_fprop_umma2_kernel<__half>;
    B200_CUDA(raise_dyn_smem_cap(kern));
    kern<<<grid, 288, smem, st>>>((const __half*)x.data, (const __half*)w, bias, (__half*)y.data, p);
  }
  B200_LAUNCH_CHECK();
  return B200_OK;
}

static int conv_fprop_umma_impl(const ActView& xv, const void* w, const float* bias, const ActView& yv_in, int kd, int kh, int kw,
                                int accumulate, cudaStream_t st, Upscale up, const ActView* phase_views, int nphases);

int conv_fprop_umma_v(const ActView& xv, const void* w, const float* bias, const ActView& yv_in, int kd, int kh, int kw,
                      int accumulate, cudaStream_t st, Upscale up) {
  return conv_fprop_umma_impl(xv, w, bias, yv_in, kd, kh, kw, accumulate, st, up, nullptr, 0);
}

// y[v][:] (+)= sum_t phase_t[v][:] * W_t^T with W packed as [t][y.c][x.c]: all phases in one launch
int conv_fprop_phases_v(const ActView* phase_views, int nphases, const void* w, const ActView& y, int accumulate, cudaStream_t st) {
  B200_CHECK_ARG(nphases >= 1 && nphases <= 8, "conv_fprop(phases): 1..8 phases");
  return conv_fprop_umma_impl(phase_views[0], w, nullptr, y, nphases, 1, 1, accumulate, st, Upscale{}, phase_views, nphases);
}

static int conv_fprop_umma_impl(const ActView& xv, const void* w, const float* bias, const ActView& yv_in, int kd, int kh, int kw,
                                int accumulate, cudaStream_t st, Upscale up, const ActView* phase_views, int nphases) {
  if (loader_mode() == 1 && up.sd == 0 && nphases == 0) return conv_fprop_umma2_v(xv, w, bias, yv_in, kd, kh, kw, accumulate, st);
  const ActView* x = &xv;
  // in transposed mode the GEMM has taps*Cout columns; `yv` keeps the fine tensor's strides and gets c = taps*Cout
  ActView yv = yv_in;
  if (up.sd) yv.c = up.sd * up.sh * up.sw * up.cout;
  const ActView* y = &yv;
  B200_CHECK_ARG(aligned16(w), "conv_fprop(umma): packed weights must be 16-byte aligned");
  FpropParams p{};
  p.n = x->n; p.d = x->d; p.h = x->h; p.w = x->w; p.cin = x->c; p.cout = y->c;
  p.kd = kd; p.kh = kh; p.kw = kw;
  pick_tile(x->d, x->h, x->w, &p.bd, &p.bh, &p.bw);
  p.ck = (x->c % 64 == 0) ? 64 : ((x->c % 32 == 0) ? 32 : 16);
  p.chunks = x->c / p.ck;
  p.nt = y->c <= 256 ? y->c : 128;
  p.tiles_d = (int)ceil_div(x->d, p.bd); p.tiles_h = (int)ceil_div(x->h, p.bh); p.tiles_w = (int)ceil_div(x->w, p.bw);
  p.tiles_n = (int)ceil_div(y->c, p.nt);
  p.num_tiles = x->n * p.tiles_d * p.tiles_h * p.tiles_w * p.tiles_n;
  p.a_bytes = 128u * p.ck * 2;
  p.b_bytes = (uint32_t)p.nt * p.ck * 2;
  p.stage_bytes = (p.a_bytes + p.b_bytes + 1023u) & ~1023u;
  int stages = (int)((200u * 1024u) / p.stage_bytes);
  if (stages > 8) stages = 8;
  B200_CHECK_ARG(stages >= 2, "conv_fprop(umma): tile does not fit in shared memory");
  p.stages = stages;
  p.layout = p.ck == 64 ? kSwizzle128 : (p.ck == 32 ? kSwizzle64 : kSwizzle32);
  p.sbo = 8u * p.ck * 2;
  p.idesc = make_idesc(x->dtype == B200_BF16, p.nt, 0, 0);
  uint32_t cols = 32;
  while (cols < 2u * p.nt) cols <<= 1;
  p.tmem_cols = cols;
  p.ysw = y->sw; p.ysh = y->sh; p.ysd = y->sd; p.ysn = y->sn;
  p.accumulate = accumulate;
  p.usd = up.sd; p.ush = up.sh; p.usw = up.sw; p.pcout = up.cout;
  p.phases = nphases; p.wrows = y->c;

  CUtensorMap tx, tw;
  PhaseMaps pm;
  memset(&pm, 0, sizeof(pm));
  int rc = make_act_tmap(&tx, *x, p.ck, p.bw, p.bh, p.bd);
  if (rc) return rc;
  // transposed mode: TMA-store epilogue through one output map per phase (B200_CONVT_TMA=0: direct scatter stores)
  static const int convt_tma_env = getenv("B200_CONVT_TMA") ? atoi(getenv("B200_CONVT_TMA")) : 1;
  const int up_phases = up.sd * up.sh * up.sw;
  if (up.sd && convt_tma_env && up_phases <= 8 && up.cout % 16 == 0 && !accumulate && aligned16(yv_in.data) && yv_in.sw % 8 == 0 &&
      yv_in.sh % 8 == 0 && yv_in.sd % 8 == 0 && yv_in.sn % 8 == 0) {
    EncodeTiledFn fn = encode_tiled_fn();
    int t = 0;
    bool ok = fn != nullptr;
    // widest box the channel count allows: fewer, longer rows per bulk store (B200_CONVT_BOX caps it: 16 / 32 / 64)
    static const int box_env = getenv("B200_CONVT_BOX") ? atoi(getenv("B200_CONVT_BOX")) : 64;
    int cbox = (up.cout % 64 == 0 && p.nt % 64 == 0) ? 64 : ((up.cout % 32 == 0 && p.nt % 32 == 0) ? 32 : 16);
    while (cbox > box_env && cbox > 16) cbox >>= 1;
    for (int a = 0; a < up.sd && ok; ++a)
      for (int b = 0; b < up.sh && ok; ++b)
        for (int c = 0; c < up.sw && ok; ++c, ++t) {
          // phase (a, b, c) of the fine tensor as a coarse tensor: x's dims, strides times the up-sampling factors
          char* base = (char*)yv_in.data + ((int64_t)a * yv_in.sd + (int64_t)b * yv_in.sh + (int64_t)c * yv_in.sw) * 2;
          const cuuint64_t dims[5] = {(cuuint64_t)up.cout, (cuuint64_t)x->w, (cuuint64_t)x->h, (cuuint64_t)x->d, (cuuint64_t)x->n};
          const cuuint64_t strides[4] = {(cuuint64_t)yv_in.sw * up.sw * 2, (cuuint64_t)yv_in.sh * up.sh * 2,
                                         (cuuint64_t)yv_in.sd * up.sd * 2, (cuuint64_t)yv_in.sn * 2};
          const cuuint32_t box[5] = {(cuuint32_t)cbox, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bd, 1};
          const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
          CUresult r = fn(&pm.m[t], tm_dtype(yv_in.dtype), 5, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          cbox == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (cbox == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          ok = r == CUDA_SUCCESS;
        }
    const uint32_t epi_bytes = 2u * (uint32_t)p.nt * 256u;       // two staging buffers of nt / 16 boxes of 4 KB
    if (ok && epi_bytes + 2u * p.stage_bytes <= 200u * 1024u) {
      int st2 = (int)((200u * 1024u - epi_bytes) / p.stage_bytes);
      if (st2 > 8) st2 = 8;
      p.stages = st2;
      p.epi_tma = 1;
      p.epi_off = (uint32_t)p.stages * p.stage_bytes;
      p.epi_cbox = cbox;
    }
  }
  if (nphases) {
    for (int t = 0; t < nphases; ++t) {
      rc = make_act_tmap(&pm.m[t], phase_views[t], p.ck, p.bw, p.bh, p.bd);
      if (rc) return rc;
    }
    rc = make_matrix_tmap(&tw, w, x->dtype, (int64_t)nphases * y->c, x->c, p.nt, p.ck);
  } else {
    rc = make_matrix_tmap(&tw, w, x->dtype, y->c, (int64_t)kd * kh * kw * x->c, p.nt, p.ck);
  }
  if (rc) return rc;

  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024 + (p.epi_tma ? 2u * (size_t)p.nt * 256u : 0);
  int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  if (x->dtype == B200_BF16) {
    auto kern = conv_fprop_umma_kernel<__nv_bfloat16>;
    B200_CUDA(raise_dyn_smem_cap(kern));
    kern<<<grid, 224, smem, st>>>(tx, tw, pm, bias, (__nv_bfloat16*)y->data, p);
  } else {
    auto kern = conv_fprop_umma_kernel<__half>;
    B200_CUDA(raise_dyn_smem_cap(kern));
    kern<<<grid, 224, smem, st>>>(tx, tw, pm, bias, (__half*)y->data, p);
  }
  B200_LAUNCH_CHECK();
  return B200_OK;
}
}  // namespace sm100

namespace sm100 {

static int make_tmap4(CUtensorMap* out, const void* base, int dtype, const cuuint64_t dims[4], const cuuint64_t strides_b[3],
                      const cuuint32_t box[4]) {
  EncodeTiledFn fn = encode_tiled_fn();
  B200_CHECK_ARG(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, tm_dtype(dtype), 4, const_cast<void*>(base), dims, strides_b, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_for_bytes((int)box[0] * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(rank 4) failed with %d", (int)r);
  return B200_OK;
}

bool conv_xfold_ok(const ActView& x, const ActView& y, int kd, int kh, int kw) {
  if (x.dtype != B200_BF16 && x.dtype != B200_F16) return false;
  if ((kw != 3 && kw != 1) || kd > 5 || kh > 5) return false;
  if (y.c % 16 != 0 || y.c > 64 || x.c > 96) return false;
  if (x.c % 16 != 0) {
    // image-fed layers (Cin = 2, 4, 8): only the unified-stage slab kernel knows the padded window
    int xoff, kxp;
    if (!xfold_geom(x.c, kw, &xoff, &kxp)) return false;
    if (kd != 3 || x.d < 8 || x.h < 16 || y.c > 16) return false;
  }
  if (x.sw != x.c || x.sh != (int64_t)x.w * x.c) return false;              // dense rows: 6 voxels are contiguous
  if (x.w % 4 != 0 || x.w < 8) return false;
  if (y.sw % 8 != 0 || !aligned16(x.data) || !aligned16(y.data)) return false;
  if ((int64_t)x.d * x.h < 128) return false;
  return encode_tiled_fn() != nullptr;
}

static int xslab_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("B200_XSLAB");
    mode = (e && strcmp(e, "0") == 0) ? 0 : 1;
  }
  return mode;
}

// The fused channel statistics are offered for Cout = 16 only.  Measured on the B200 (tools/epi_micro.py, @128^3 / @64^3 x4):
// +0.027 ms on 16 -> 16 against 0.061 ms for b200_channel_sums, but +0.032 ms on 32 -> 32 @64^3 (stand-alone: 0.022 ms) and
// +0.27 ms on 16 -> 48 -- 64 / 96 running sums per thread push the epilogue warps past the 168-register budget.
static bool xslab_stats_ok(const ActView& y, const double* stats) {
  if (!stats) return false;
  if (y.c != 16) return false;
  return getenv("B200_NO_EPI_STATS") == nullptr;
}

static int conv_fprop_xslab_v(const ActView& x, const void* w, const float* bias, const ActView& y, int kd, int kh, int kw,
                              int accumulate, cudaStream_t st, double* stats = nullptr, int* stats_applied = nullptr) {
  XslabParams p{};
  // TMA-store epilogue (default): needs a 16-byte aligned output whose voxel pitch is a multiple of 16 bytes.  B200_EPI_TMA=0
  // selects the direct-store epilogue, B200_EPI_SWZ = 128 (default) / 64 / 32 / 0 the swizzle of the staging buffers.
  static const int epi_env = getenv("B200_EPI_TMA") ? atoi(getenv("B200_EPI_TMA")) : 1;
  // (SWIZZLE_128B with the 32-byte inner box faults on the B200 -- illegal memory access; 32B is the box's own span)
  static const int epi_swz_env = getenv("B200_EPI_SWZ") ? atoi(getenv("B200_EPI_SWZ")) : 32;
  p.epi_tma = (epi_env != 0 && y.c % 16 == 0 && y.sw % 8 == 0 && y.sh % 8 == 0 && y.sd % 8 == 0 && y.sn % 8 == 0 && aligned16(y.data)) ? 1 : 0;
  p.epi_swz = epi_swz_env == 128 ? 7u : (epi_swz_env == 64 ? 3u : (epi_swz_env == 32 ? 1u : 0u));
  // the element-wise add of the TMA unit never shows the final value to the SM: no fused reduction on accumulating launches
  const int cb = (xslab_stats_ok(y, stats) && !(p.epi_tma && accumulate)) ? y.c / 16 : 0;
  p.stats = cb ? stats : nullptr;
  if (stats_applied) *stats_applied = cb ? 1 : 0;
  const uint32_t ring_budget = p.epi_tma ? 188u * 1024u : 200u * 1024u;
  p.n = x.n; p.d = x.d; p.h = x.h; p.w = x.w; p.cin = x.c; p.cout = y.c;
  p.kd = kd; p.kh = kh; p.kw = kw;
  p.nt = 4 * y.c;
  p.rt = (x.d >= 16) ? 2 : 1;
  // Resident weights (conv_fprop_xslab1_kernel, p.wres): when the Toeplitz matrix of the layer (nt x kd*kh*kx elements: 108 KB
  // for 16 -> 16, 36 KB for 2 -> 16) fits next to the slab ring it is fetched ONCE per CTA instead of once per tile.  The x-slab
  // kernels run at the L2 -> shared-memory rate of the TMA unit (~70 B/clk/SM, tools/pipe_rates.py): 16 -> 16 moves 270 KB per
  // 512 voxels, 108 KB of it weights.  Row tiles per CTA tile drop to 1 when only two 2-row-tile slabs would fit.
  // (Also measured and dropped: an extra warp issuing cp.async.bulk.prefetch.tensor 1 / 2 / 4 tiles ahead of the slab producer.
  // Every x-slab layer got slower -- 16 -> 16 0.252 -> 0.35 ms, 48 -> 16 0.60 -> 0.71 ms, step 15.7 -> 16.3 ms: the prefetches
  // go through the same TMA unit and take request slots from the loads.  profiles/xslab_negative_results_r1.log)
  static const int wres_env = getenv("B200_WRES") ? atoi(getenv("B200_WRES")) : 1;
  static const int wres_rt_env = getenv("B200_WRES_RT") ? atoi(getenv("B200_WRES_RT")) : 0;
  uint32_t wres_bytes = 0;
  {
    int xo = 0, kxp = 0;
    if (wres_env && kd <= 3 && xfold_geom(x.c, kw, &xo, &kxp)) {
      const uint32_t wb = (uint32_t)(4 * y.c) * (uint32_t)(kd * kh * kxp) * 2u;
      const uint32_t avail = 232448u - 1024u - 1024u - (p.epi_tma ? 32768u : 0u);     // 227 KB - static - alignment - staging
      if (wb <= 112u * 1024u) {
        const uint32_t slab2 = (uint32_t)(16 + kd - 1) * 16u * 128u, slab1 = (uint32_t)(8 + kd - 1) * 16u * 128u;
        int rt = p.rt;
        if (wres_rt_env == 1 || wres_rt_env == 2) rt = (wres_rt_env == 2 && x.d >= 16) ? 2 : 1;
        else if (rt == 2 && (avail - wb) / slab2 < 3) rt = 1;
        // Default (B200_WRES=1): only when the ring keeps >= 4 slabs in flight -- the kernels run at bytes-in-flight / latency,
        // and 16 -> 16 lost with 2 x 36 KB or 4 x 20 KB slabs (0.251 -> 0.30 ms) while 2 -> 16 won (0.191 -> 0.131 ms).
        // B200_WRES=2 takes every layer whose matrix fits.
        const uint32_t slabs = (avail - wb) / (rt == 2 ? slab2 : slab1);
        if (slabs >= (wres_env >= 2 ? 2u : 4u) && (wres_env >= 2 || rt == p.rt)) {
          wres_bytes = wb;
          p.rt = rt;
        }
      }
    }
  }
  p.nbuf = (2 * p.rt * p.nt <= 512) ? 2 : 1;
  p.zl = 8 * p.rt + kd - 1;
  p.groups_x = x.w / 4;
  p.tiles_h = (int)ceil_div(x.h, 16); p.tiles_d = (int)ceil_div(x.d, 8 * p.rt);
  p.num_tiles = x.n * p.tiles_d * p.tiles_h * p.groups_x;
  B200_CHECK_ARG(xfold_geom(x.c, kw, &p.xoff, &p.kx), "conv_fprop(xslab): unsupported Cin");
  p.boxes64 = p.kx / 64; p.has32 = (p.kx % 64) ? 1 : 0;           // has32 = number of 32-element boxes behind the 64-element ones
  p.a_bytes = (((uint32_t)p.zl * 16u * 128u) + 1023u) & ~1023u;
  p.b_bytes = (((uint32_t)p.nt * 128u) + 1023u) & ~1023u;
  // split ~200 KB between the two rings: at least 2 slabs, the rest for weight tiles (3..6)
  int a_st = 3, b_st;
  for (;;) {
    b_st = (int)((ring_budget - (uint32_t)a_st * p.a_bytes) / p.b_bytes);
    if (b_st >= 3 || a_st == 2) break;
    --a_st;
  }
  if (b_st > 6) b_st = 6;
  B200_CHECK_ARG(b_st >= 2, "conv_fprop(xslab): tiles do not fit in shared memory");
  p.a_stages = a_st; p.b_stages = b_st;
  p.b_off = (uint32_t)a_st * p.a_bytes;
  p.idesc = make_idesc(x.dtype == B200_BF16, p.nt, 0, 0);
  uint32_t cols = 32;
  while (cols < (uint32_t)(p.nbuf * p.rt * p.nt)) cols <<= 1;
  p.tmem_cols = cols;
  p.ysw = y.sw; p.ysh = y.sh; p.ysd = y.sd; p.ysn = y.sn;
  p.accumulate = accumulate;
  {
    const char* e = getenv("B200_ABLATE");
    p.ablate = e ? atoi(e) : 0;
  }
  static long long* dbg_buf = nullptr;
  if (getenv("B200_DBG")) {
    if (!dbg_buf) B200_CUDA(cudaMalloc(&dbg_buf, sizeof(long long) * 16 * 148));
    B200_CUDA(cudaMemsetAsync(dbg_buf, 0, sizeof(long long) * 16 * 148, st));
    p.dbg = dbg_buf;
  }

  CUtensorMap tx64, tx32, tw64, tw32;
  const cuuint64_t xd[4] = {(cuuint64_t)x.w * x.c, (cuuint64_t)x.h, (cuuint64_t)x.d, (cuuint64_t)x.n};
  const cuuint64_t xs[3] = {(cuuint64_t)x.sh * 2, (cuuint64_t)x.sd * 2, (cuuint64_t)x.sn * 2};
  const cuuint32_t b64[4] = {64, 16, (cuuint32_t)p.zl, 1};
  const cuuint32_t b32[4] = {32, 16, (cuuint32_t)p.zl, 1};
  int rc = make_tmap4(&tx64, x.data, x.dtype, xd, xs, b64);
  if (rc) return rc;
  rc = make_tmap4(&tx32, x.data, x.dtype, xd, xs, b32);
  if (rc) return rc;
  const int64_t ktot = (int64_t)kd * kh * p.kx;
  rc = make_matrix_tmap(&tw64, w, x.dtype, p.nt, ktot, p.nt, 64);
  if (rc) return rc;
  rc = make_matrix_tmap(&tw32, w, x.dtype, p.nt, ktot, p.nt, 32);
  if (rc) return rc;

  // output map of the TMA-store epilogue: (C, W, H, D, N), box = 16 channels x 4 voxels x 16 rows x 8 lines
  CUtensorMap ty = tx64;
  if (p.epi_tma) {
    EncodeTiledFn fn = encode_tiled_fn();
    const cuuint64_t yd[5] = {(cuuint64_t)y.c, (cuuint64_t)y.w, (cuuint64_t)y.h, (cuuint64_t)y.d, (cuuint64_t)y.n};
    const cuuint64_t ys[4] = {(cuuint64_t)y.sw * 2, (cuuint64_t)y.sh * 2, (cuuint64_t)y.sd * 2, (cuuint64_t)y.sn * 2};
    const cuuint32_t yb[5] = {16, 4, 16, 8, 1};
    const cuuint32_t ye[5] = {1, 1, 1, 1, 1};
    const CUtensorMapSwizzle sw = p.epi_swz == 7u ? CU_TENSOR_MAP_SWIZZLE_128B : (p.epi_swz == 3u ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : (p.epi_swz == 1u ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE));
    CUresult r = fn(&ty, tm_dtype(y.dtype), 5, y.data, yd, ys, yb, ye, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(xslab output) failed with %d (c=%d sw=%lld)", (int)r, y.c, (long long)y.sw);
  }

  int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  // narrow layers: slab box + its kd weight tiles in ONE stage (conv_fprop_xslab1_kernel) when >= 3 such stages fit
  // A 96-element window (Cin = 16, kw = 3) is one 64-box + one 32-box.  Cut into three 32-boxes a unified stage halves, so the
  // N = 128 / 192 layers (16 -> 32, 16 -> 48) fit the unified-stage kernel, which they do not with 128-byte rows: 16 -> 48 @128^3
  // 0.729 -> 0.621 ms (9 stage hand-overs per tile instead of 24).  Where the 128-byte stages fit anyway (16 -> 16) the cut only
  // adds hand-overs (6 -> 9 per tile): 0.251 -> 0.273 ms, so it is not applied there (B200_XSLAB_ALL32: 0 never, 1 always,
  // 2 = default, only when needed).
  static const int all32_env = getenv("B200_XSLAB_ALL32") ? atoi(getenv("B200_XSLAB_ALL32")) : 2;
  static const bool allow_unified = !(getenv("B200_XSLAB_UNIFIED") && strcmp(getenv("B200_XSLAB_UNIFIED"), "0") == 0);
  uint32_t uni_bytes = (uint32_t)p.zl * 16u * 128u + (uint32_t)kd * (uint32_t)p.nt * 128u;
  if (all32_env && p.kx == 96 && kd <= 3 && !wres_bytes && allow_unified && !p.ablate && !p.dbg &&
      (all32_env == 1 || 3u * uni_bytes > ring_budget)) {
    const uint32_t u32 = (uint32_t)p.zl * 16u * 64u + (uint32_t)kd * (uint32_t)p.nt * 64u;
    if (3u * u32 <= ring_budget) { p.boxes64 = 0; p.has32 = 3; uni_bytes = u32; }
  }
  if (wres_bytes && !p.ablate && !p.dbg) {
    const uint32_t avail = 232448u - 1024u - 1024u - (p.epi_tma ? 32768u : 0u);
    p.wres = 1;
    p.a_bytes = (uint32_t)p.zl * 16u * 128u;                       // slab box only (a multiple of 2 KB)
    p.a_stages = (int)((avail - wres_bytes) / p.a_bytes);
    if (p.a_stages > 6) p.a_stages = 6;
    p.wres_off = (uint32_t)p.a_stages * p.a_bytes;
    p.epi_off = (p.wres_off + wres_bytes + 1023u) & ~1023u;
  }
  if (p.wres || (allow_unified && kd <= 3 && 3u * uni_bytes <= ring_budget && !p.ablate && !p.dbg)) {
    if (!p.wres) {
      p.a_bytes = uni_bytes;
      p.a_stages = (int)(ring_budget / uni_bytes);
      if (p.a_stages > 6) p.a_stages = 6;
      p.epi_off = ((uint32_t)p.a_stages * p.a_bytes + 1023u) & ~1023u;
      // 64-box + 32-box windows (Cin = 16): the 32-box stage needs half a slot.  Alternating big / small slots puts 4 stages
      // where 3 uniform ones fit -- same hand-overs per tile, one more stage of TMA latency covered.
      static const int vs_env = getenv("B200_XSLAB_VARSLOT") ? atoi(getenv("B200_XSLAB_VARSLOT")) : 1;
      const uint32_t pair = uni_bytes + uni_bytes / 2;
      if (vs_env && p.boxes64 == 1 && p.has32 == 1 && p.a_stages < 4 && 2u * pair <= ring_budget && (uni_bytes / 2) % 1024u == 0) {
        p.varslot = 1;
        p.a_stages = 4;
        p.epi_off = (2u * pair + 1023u) & ~1023u;
      }
    }
    const size_t smem1 = (size_t)p.epi_off + (p.epi_tma ? 32768u : 0u) + 1024;
#define XSLAB1_LAUNCH(TT, CBV)                                                                            \
  {                                                                                                      \
    auto kern = conv_fprop_xslab1_kernel<TT, CBV>;                                                       \
    B200_CUDA(raise_dyn_smem_cap(kern));                                                                 \
    kern<<<grid, 320, smem1, st>>>(tx64, tx32, tw64, tw32, ty, bias, (TT*)y.data, p);                        \
  }
#define XSLAB1_BY_CB(TT)                                                                                 \
  switch (cb) {                                                                                          \
    case 1: XSLAB1_LAUNCH(TT, 1) break;                                                                  \
    default: XSLAB1_LAUNCH(TT, 0) break;                                                                 \
  }
    if (x.dtype == B200_BF16) XSLAB1_BY_CB(__nv_bfloat16) else XSLAB1_BY_CB(__half)
#undef XSLAB1_BY_CB
#undef XSLAB1_LAUNCH
    B200_LAUNCH_CHECK();
    return B200_OK;
  }
  B200_CHECK_ARG(p.xoff == 0, "conv_fprop(xslab): image-fed layers need the unified-stage kernel");
  p.epi_off = (p.b_off + (uint32_t)p.b_stages * p.b_bytes + 1023u) & ~1023u;
  const size_t smem = (size_t)p.epi_off + (p.epi_tma ? 32768u : 0u) + 1024;
  const int threads = 32 * (6 + kSlabAWarps + kSlabBWarps);
#define XSLAB_LAUNCH(TT, CBV)                                                                             \
  {                                                                                                      \
    auto kern = conv_fprop_xslab_kernel<TT, CBV>;                                                        \
    B200_CUDA(raise_dyn_smem_cap(kern));                                                                 \
    kern<<<grid, threads, smem, st>>>(tx64, tx32, tw64, tw32, ty, bias, (TT*)y.data, p);                     \
  }
#define XSLAB_BY_CB(TT)                                                                                  \
  switch (cb) {                                                                                          \
    case 1: XSLAB_LAUNCH(TT, 1) break;                                                                   \
    default: XSLAB_LAUNCH(TT, 0) break;                                                                  \
  }
  if (x.dtype == B200_BF16) XSLAB_BY_CB(__nv_bfloat16) else XSLAB_BY_CB(__half)
#undef XSLAB_BY_CB
#undef XSLAB_LAUNCH
  B200_LAUNCH_CHECK();
  if (p.dbg) {
    long long h[16 * 148];
    B200_CUDA(cudaMemcpyAsync(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    double tot[2] = {0, 0}, te[2] = {0, 0}, af[2] = {0, 0}, bf[2] = {0, 0}, ep = 0;
    for (int b = 0; b < grid; ++b) {
      for (int r = 0; r < 2; ++r) {
        tot[r] += h[b * 16 + r * 8]; te[r] += h[b * 16 + r * 8 + 1]; af[r] += h[b * 16 + r * 8 + 2]; bf[r] += h[b * 16 + r * 8 + 3];
      }
      ep += h[b * 16 + 4];
    }
    const double tiles = (double)p.num_tiles;
    printf("xslab dbg: per CTA tile  MMA warp0: loop %.0f  wait tmem-empty %.0f  wait slab %.0f  wait weights %.0f | warp1: loop %.0f "
           "tmem %.0f slab %.0f weights %.0f | epilogue wait tmem-full %.0f\n",
           tot[0] / tiles, te[0] / tiles, af[0] / tiles, bf[0] / tiles, tot[1] / tiles, te[1] / tiles, af[1] / tiles, bf[1] / tiles,
           ep / tiles);
    fflush(stdout);
  }
  return B200_OK;
}

int conv_fprop_xfold_v(const ActView& x, const void* w, const float* bias, const ActView& y, int kd, int kh, int kw,
                       int accumulate, cudaStream_t st, double* stats = nullptr, int* stats_applied = nullptr) {
  B200_CHECK_ARG(conv_xfold_ok(x, y, kd, kh, kw), "conv_fprop(xfold): unsupported operands");
  if (stats_applied) *stats_applied = 0;
  if (xslab_mode() && kd == 3 && x.d >= 8 && x.h >= 16)
    return conv_fprop_xslab_v(x, w, bias, y, kd, kh, kw, accumulate, st, stats, stats_applied);
  XfoldParams p{};
  p.n = x.n; p.d = x.d; p.h = x.h; p.w = x.w; p.cin = x.c; p.cout = y.c;
  p.kd = kd; p.kh = kh; p.kw = kw;
  {  // rows of a tile = bh x bd positions in (y, z)
    static const int cand[][2] = {{16, 8}, {8, 16}, {32, 4}, {4, 32}, {64, 2}, {2, 64}, {128, 1}, {1, 128}};
    double best = 1e30;
    for (auto& c : cand) {
      double cover = (double)ceil_div(x.h, c[0]) * c[0] * ceil_div(x.d, c[1]) * c[1];
      if (cover < best - 0.5) { best = cover; p.bh = c[0]; p.bd = c[1]; }
    }
  }
  p.groups_x = x.w / 4;
  p.tiles_h = (int)ceil_div(x.h, p.bh); p.tiles_d = (int)ceil_div(x.d, p.bd);
  p.num_tiles = x.n * p.tiles_d * p.tiles_h * p.groups_x;
  p.kx = (3 + kw) * x.c;
  p.boxes64 = p.kx / 64; p.has32 = (p.kx % 64) ? 1 : 0;
  p.nt = 4 * y.c;
  p.a_bytes = 128u * 128u;
  p.stage_bytes = p.a_bytes + (((uint32_t)p.nt * 128u + 1023u) & ~1023u);
  int stages = (int)((200u * 1024u) / p.stage_bytes);
  if (stages > 8) stages = 8;
  B200_CHECK_ARG(stages >= 2, "conv_fprop(xfold): tile does not fit in shared memory");
  p.stages = stages;
  p.idesc = make_idesc(x.dtype == B200_BF16, p.nt, 0, 0);
  uint32_t cols = 32;
  while (cols < 2u * p.nt) cols <<= 1;
  p.tmem_cols = cols;
  p.ysw = y.sw; p.ysh = y.sh; p.ysd = y.sd; p.ysn = y.sn;
  p.accumulate = accumulate;

  CUtensorMap tx64, tx32, tw64, tw32;
  const cuuint64_t xd[4] = {(cuuint64_t)x.w * x.c, (cuuint64_t)x.h, (cuuint64_t)x.d, (cuuint64_t)x.n};
  const cuuint64_t xs[3] = {(cuuint64_t)x.sh * 2, (cuuint64_t)x.sd * 2, (cuuint64_t)x.sn * 2};
  const cuuint32_t b64[4] = {64, (cuuint32_t)p.bh, (cuuint32_t)p.bd, 1};
  const cuuint32_t b32[4] = {32, (cuuint32_t)p.bh, (cuuint32_t)p.bd, 1};
  int rc = make_tmap4(&tx64, x.data, x.dtype, xd, xs, b64);
  if (rc) return rc;
  rc = make_tmap4(&tx32, x.data, x.dtype, xd, xs, b32);
  if (rc) return rc;
  const int64_t ktot = (int64_t)kd * kh * p.kx;
  rc = make_matrix_tmap(&tw64, w, x.dtype, p.nt, ktot, p.nt, 64);
  if (rc) return rc;
  rc = make_matrix_tmap(&tw32, w, x.dtype, p.nt, ktot, p.nt, 32);
  if (rc) return rc;

  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
  int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  if (x.dtype == B200_BF16) {
    auto kern = conv_fprop_xfold_kernel<__nv_bfloat16>;
    B200_CUDA(raise_dyn_smem_cap(kern));
    kern<<<grid, 224, smem, st>>>(tx64, tx32, tw64, tw32, bias, (__nv_bfloat16*)y.data, p);
  } else {
    auto kern = conv_fprop_xfold_kernel<__half>;
    B200_CUDA(raise_dyn_smem_cap(kern));
    kern<<<grid, 224, smem, st>>>(tx64, tx32, tw64, tw32, bias, (__half*)y.data, p);
  }
  B200_LAUNCH_CHECK();
  return B200_OK;
}

}  // namespace sm100

bool conv_fprop_xfold_supported(const b200_tensor* x, const b200_tensor* y, int kd, int kh, int kw) {
  return sm100::conv_xfold_ok(sm100::view_of(x), sm100::view_of(y), kd, kh, kw);
}
int conv_fprop_xfold(const b200_tensor* x, const void* w, const float* bias, const b200_tensor* y, int kd, int kh, int kw,
                     int accumulate, cudaStream_t st) {
  return sm100::conv_fprop_xfold_v(sm100::view_of(x), w, bias, sm100::view_of(y), kd, kh, kw, accumulate, st);
}
}  // namespace b200

B200_EXPORT int b200_conv_fprop_stats(const b200_tensor* x, const void* w_packed, const float* bias, const b200_tensor* y,
                                      int32_t kd, int32_t kh, int32_t kw, int32_t accumulate, double* sums, int32_t* applied,
                                      void* stream) {
  using namespace b200;
  B200_CHECK_ARG(x && y && w_packed && sums && applied, "conv_fprop_stats: null pointer");
  B200_CHECK_ARG(check_tensor(x, "conv_fprop_stats.x") && check_tensor(y, "conv_fprop_stats.y"), "%s", b200_last_error());
  B200_CHECK_ARG(conv_fprop_xfold_supported(x, y, kd, kh, kw),
                 "conv_fprop_stats: operands not supported by the x-folded kernel (query b200_conv_impl_query)");
  int ok = 0;
  int rc = sm100::conv_fprop_xfold_v(sm100::view_of(x), w_packed, bias, sm100::view_of(y), kd, kh, kw, accumulate, (cudaStream_t)stream,
                                     sums, &ok);
  *applied = ok;
  return rc;
}

namespace b200 {
int pack_weight_xfold(const float* w, void* packed, int dtype, int cout, int cin, int kd, int kh, int kw, int flip,
                      cudaStream_t st) {
  const int CO = flip ? cin : cout, CI = flip ? cout : cin;
  int xoff = 0, kxp = 0;
  B200_CHECK_ARG(sm100::xfold_geom(CI, kw, &xoff, &kxp), "pack_weight_xfold: unsupported input channel count %d", CI);
  int64_t total = (int64_t)4 * CO * kd * kh * kxp;
  unsigned blocks = (unsigned)(ceil_div(total, 256) < 8192 ? ceil_div(total, 256) : 8192);
  if (dtype == B200_BF16)
    sm100::pack_weight_xfold_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(w, (__nv_bfloat16*)packed, cout, cin, kd, kh, kw, flip);
  else if (dtype == B200_F16)
    sm100::pack_weight_xfold_kernel<__half><<<blocks, 256, 0, st>>>(w, (__half*)packed, cout, cin, kd, kh, kw, flip);
  else {
    set_error("pack_weight_xfold: 16-bit dtypes only");
    return B200_ERR_ARG;
  }
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// ================================================================================================== wgrad
// dW[co][tap][ci] += sum_vox dY[vox][co] * X[vox + off(tap)][ci]  as  D[(tap, ci)][co] = A^T B with K = voxels:
//   A  = shifted activation tiles [128 voxels][16 ch]  (TMA, 32-byte rows, SWIZZLE_32B)  -> M-major operand
//   B  = gradient tile           [128 voxels][Cout]    (TMA, 32/64/128-byte rows)        -> N-major operand
// The same bytes the fprop kernel reads K-major are consumed here MN-major (instruction-descriptor major bits = 1):
// channels are the contiguous dimension and the voxel index is K.  Eight (tap, 16-channel) chunks form one M = 128
// block (LBO = 4 KB chunk tile, SBO = 256 B per 8 voxels); each CTA owns `g` M-blocks whose fp32 accumulators stay
// in TMEM (g * Cout <= 512 columns) across ALL voxel tiles it visits, and are added to dw with fp32 atomics once.
namespace sm100 {

struct WgradParams {
  int n, d, h, w, cin, cout;
  int kd, kh, kw;
  int bd, bh, bw;
  int tiles_d, tiles_h, tiles_w, num_vtiles;
  int chunks16;            // cin / 16
  int q_total;             // taps * chunks16
  int mb_total;            // ceil(q_total / 8)
  int g;                   // M-blocks per CTA
  int a_stages, b_stages;
  uint32_t b_bytes, b_box_c, b_boxes, b_layout, b_sbo, b_lbo, b_kstep;
  uint32_t idesc, tmem_cols;
  uint32_t a_off;          // byte offset of the A ring inside dynamic smem (after the B ring)
  int bpp;                 // > 0: phase mode -- B box i is channel box (i % bpp) of PhaseMaps::m[i / bpp]
};

constexpr uint32_t kChunkBytes = 128u * 16u * 2u;   // one [128 voxels][16 ch] tile
constexpr uint32_t kBlockBytes = 8u * kChunkBytes;   // one M-block of A
constexpr int kMaxAStages = 6, kMaxBStages = 6;

template <typename T>
__global__ void __launch_bounds__(320, 1)
conv_wgrad_umma_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                       const __grid_constant__ PhaseMaps pm, float* __restrict__ dw, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_bar[2 * kMaxAStages + 2 * kMaxBStages + 1];
  __shared__ uint32_t s_tmem;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_afull = smem_u32(&s_bar[0]);
  const uint32_t bar_aempty = smem_u32(&s_bar[kMaxAStages]);
  const uint32_t bar_bfull = smem_u32(&s_bar[2 * kMaxAStages]);
  const uint32_t bar_bempty = smem_u32(&s_bar[2 * kMaxAStages + kMaxBStages]);
  const uint32_t bar_done = smem_u32(&s_bar[2 * kMaxAStages + 2 * kMaxBStages]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(bar_afull + 8 * s, 4); mbar_init(bar_aempty + 8 * s, 1); }   // 4 chunk producers
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(bar_bfull + 8 * s, 1); mbar_init(bar_bempty + 8 * s, 1); }
    mbar_init(bar_done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_dy);
  }
  if (warp == 1) tmem_alloc(smem_u32(&s_tmem), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  const int pd = p.kd / 2, ph = p.kh / 2, pw = p.kw / 2;
  const int mb0 = blockIdx.y * p.g;
  const int mb1 = min(mb0 + p.g, p.mb_total);

  auto decode = [&](int vt, int& n, int& z0, int& y0, int& x0) {
    int t = vt;
    x0 = (t % p.tiles_w) * p.bw; t /= p.tiles_w;
    y0 = (t % p.tiles_h) * p.bh; t /= p.tiles_h;
    z0 = (t % p.tiles_d) * p.bd; t /= p.tiles_d;
    n = t;
  };

  if (warp == 0) {
    if (elect_one()) {
      // ================================================================= TMA producer of the dY tiles (B operand)
      int bs = 0;
      uint32_t bph = 0;
      for (int vt = blockIdx.x; vt < p.num_vtiles; vt += gridDim.x) {
        int n, z0, y0, x0;
        decode(vt, n, z0, y0, x0);
        mbar_wait(bar_bempty + 8 * bs, bph ^ 1);
        const uint32_t b_dst = smem0 + bs * p.b_bytes;
        mbar_expect_tx(bar_bfull + 8 * bs, p.b_bytes);
        if (p.bpp > 0) {
          uint32_t i = 0;
          for (int t = 0; i < p.b_boxes; ++t)
            for (int c = 0; c < p.bpp; ++c, ++i)
              tma_load_5d(b_dst + i * 128u * p.b_box_c * 2u, &pm.m[t], bar_bfull + 8 * bs, (int)(c * p.b_box_c), x0, y0, z0, n);
        } else {
          for (uint32_t i = 0; i < p.b_boxes; ++i)
            tma_load_5d(b_dst + i * 128u * p.b_box_c * 2u, &tmap_dy, bar_bfull + 8 * bs, (int)(i * p.b_box_c), x0, y0, z0, n);
        }
        if (++bs == p.b_stages) { bs = 0; bph ^= 1; }
      }
    }
  } else if (warp >= 6) {
    if (elect_one()) {
      // ================================================================= TMA producers of the input chunks (A operand)
      // one thread sustains one bulk-tensor copy per ~450 cycles (tools/pipe_rates.py): warp 6 + j fetches chunks j and
      // j + 4 of every M-block; all four warps arrive on the block's full barrier
      const int j = warp - 6;
      const int chunks16 = p.chunks16, kw = p.kw, kh = p.kh, a_stages = p.a_stages, q_total = p.q_total;
      struct Pos { int cc, dx, dy, dz; };
      auto advance = [&](Pos& c, int steps) {
        c.cc += steps;
        while (c.cc >= chunks16) { c.cc -= chunks16; if (++c.dx == kw) { c.dx = 0; if (++c.dy == kh) { c.dy = 0; ++c.dz; } } }
      };
      Pos start{0, 0, 0, 0};
      {
        const int q = mb0 * 8 + j;
        const int tap = q / chunks16;
        start.cc = q - tap * chunks16;
        start.dx = tap % kw; start.dy = (tap / kw) % kh; start.dz = tap / (kw * kh);
      }
      int as = 0;
      uint32_t aph = 0;
      for (int vt = blockIdx.x; vt < p.num_vtiles; vt += gridDim.x) {
        int n, z0, y0, x0;
        decode(vt, n, z0, y0, x0);
        Pos c0 = start, c1 = start;
        advance(c1, 4);
        for (int mb = mb0; mb < mb1; ++mb) {
          mbar_wait(bar_aempty + 8 * as, aph ^ 1);
          const uint32_t fa = bar_afull + 8 * as;
          const uint32_t a_dst = smem0 + p.a_off + as * kBlockBytes;
          const int q = mb * 8 + j;
          const int mine = (q < q_total ? 1 : 0) + (q + 4 < q_total ? 1 : 0);
          if (mine) {
            mbar_expect_tx(fa, (uint32_t)mine * kChunkBytes);
            tma_load_5d(a_dst + j * kChunkBytes, &tmap_x, fa, c0.cc * 16, x0 + c0.dx - pw, y0 + c0.dy - ph, z0 + c0.dz - pd, n);
            if (mine == 2)
              tma_load_5d(a_dst + (j + 4) * kChunkBytes, &tmap_x, fa, c1.cc * 16, x0 + c1.dx - pw, y0 + c1.dy - ph,
                          z0 + c1.dz - pd, n);
          } else {
            mbar_arrive(fa);
          }
          advance(c0, 8);
          advance(c1, 8);
          if (++as == a_stages) { as = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ================================================================= MMA issuer
      // A: M-major, SWIZZLE_32B: 16 voxels (K) per instruction = 512 bytes of every chunk tile
      const uint64_t tmpl_a = make_smem_desc(0, kChunkBytes, 256u, kSwizzle32);
      const uint64_t tmpl_b = make_smem_desc(0, p.b_lbo, p.b_sbo, p.b_layout);
      const uint32_t a0 = (smem0 + p.a_off) >> 4, b0 = smem0 >> 4, b16 = p.b_bytes >> 4, bstep16 = p.b_kstep >> 4;
      const uint32_t idesc = p.idesc, cout = (uint32_t)p.cout;
      const int a_stages = p.a_stages, b_stages = p.b_stages;
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0, acc = 0;
      for (int vt = blockIdx.x; vt < p.num_vtiles; vt += gridDim.x) {
        mbar_wait(bar_bfull + 8 * bs, bph);
        tc_fence_after();
        const uint64_t bd = tmpl_b + (uint64_t)(b0 + (uint32_t)bs * b16);
        uint32_t d_tmem = tmem;
        for (int mb = mb0; mb < mb1; ++mb, d_tmem += cout) {
          mbar_wait(bar_afull + 8 * as, aph);
          tc_fence_after();
          const uint64_t ad = tmpl_a + (uint64_t)(a0 + (uint32_t)as * (kBlockBytes >> 4));
          umma_ksteps_strided<8>(d_tmem, ad, bd, 512u >> 4, bstep16, idesc, acc);
          umma_commit(bar_aempty + 8 * as);
          if (++as == a_stages) { as = 0; aph ^= 1; }
        }
        umma_commit(bar_bempty + 8 * bs);
        if (++bs == b_stages) { bs = 0; bph ^= 1; }
        acc = 1;
      }
      umma_commit(bar_done);
    }
  } else {
    // =================================================================== epilogue: TMEM -> fp32 atomics into dw
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    const int taps = p.kd * p.kh * p.kw;
    if ((int)blockIdx.x < p.num_vtiles) {
      mbar_wait(bar_done, 0);
      tc_fence_after();
      for (int mb = mb0; mb < mb1; ++mb) {
        const int q = mb * 8 + row / 16;
        const bool valid = q < p.q_total;
        const int tap = valid ? q / p.chunks16 : 0;
        const int ci = valid ? (q - tap * p.chunks16) * 16 + (row & 15) : 0;
        const uint32_t taddr = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)((mb - mb0) * p.cout);
        for (int j0 = 0; j0 < p.cout; j0 += 16) {
          uint32_t r[16];
          tmem_ld16(taddr + j0, r);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              atomicAdd(dw + ((int64_t)(j0 + j) * taps + tap) * p.cin + ci, __uint_as_float(r[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------ x-folded wgrad
// Weight gradient of the small-channel layers in the x-folded formulation (see conv_fprop_xfold_kernel):
//     dWt[(dz,dy,xi,ci)][(j,co)] = sum over rows (n,z,y) and x-groups g of  X[z+dz, y+dy, 4g-pw+xi, ci] * dY[z, y, 4g+j, co]
// GEMM rows (K) are (y,z) positions; the A operand is the input window (win*Cin contiguous elements per row) fetched as
// 32-element / 64-byte TMA boxes = M-major SWIZZLE_64B atoms (4 atoms = one M = 128 block), the B operand is the dY row
// (4 voxels x Cout contiguous elements) as N-major SWIZZLE_128B atoms of 64 elements.  The epilogue folds the Toeplitz
// diagonals straight into dw[co][tap][ci] with fp32 atomics (kx = xi - j; entries outside [0, kw) are structural zeros).
struct WgradXParams {
  int n, d, h, w, cin, cout;
  int kd, kh, kw, win;
  int bh, bd, groups_x, tiles_h, tiles_d, num_vtiles;
  int atoms_per_slab;      // win * cin / 32
  int q_total;             // kd * kh * atoms_per_slab
  int mb_total, g;         // M-blocks of 4 atoms; M-blocks per CTA
  int a_stages, b_stages;
  int nt;                  // 4 * cout
  uint32_t b_boxes, b_bytes, a_off;
  uint32_t idesc, tmem_cols;
};

constexpr uint32_t kXAtomBytes = 128u * 64u;          // [128 rows][32 el]
constexpr uint32_t kXBlockBytes = 4u * kXAtomBytes;    // one M-block of A

template <typename T>
__global__ void __launch_bounds__(320, 1)
conv_wgrad_xfold_kernel(const __grid_constant__ CUtensorMap tmx32, const __grid_constant__ CUtensorMap tmy64,
                        float* __restrict__ dw, const WgradXParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_bar[2 * kMaxAStages + 2 * kMaxBStages + 1];
  __shared__ uint32_t s_tmem;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_afull = smem_u32(&s_bar[0]);
  const uint32_t bar_aempty = smem_u32(&s_bar[kMaxAStages]);
  const uint32_t bar_bfull = smem_u32(&s_bar[2 * kMaxAStages]);
  const uint32_t bar_bempty = smem_u32(&s_bar[2 * kMaxAStages + kMaxBStages]);
  const uint32_t bar_done = smem_u32(&s_bar[2 * kMaxAStages + 2 * kMaxBStages]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(bar_afull + 8 * s, 4); mbar_init(bar_aempty + 8 * s, 1); }   // 4 atom producers
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(bar_bfull + 8 * s, 1); mbar_init(bar_bempty + 8 * s, 1); }
    mbar_init(bar_done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmx32);
    tma_prefetch_desc(&tmy64);
  }
  if (warp == 1) tmem_alloc(smem_u32(&s_tmem), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  const int pd = p.kd / 2, ph = p.kh / 2, pw = p.kw / 2;
  const int mb0 = blockIdx.y * p.g;
  const int mb1 = min(mb0 + p.g, p.mb_total);

  auto decode = [&](int vt, int& n, int& z0, int& y0, int& g) {
    int t = vt;
    g = t % p.groups_x; t /= p.groups_x;
    y0 = (t % p.tiles_h) * p.bh; t /= p.tiles_h;
    z0 = (t % p.tiles_d) * p.bd; t /= p.tiles_d;
    n = t;
  };

  if (warp == 0) {
    if (elect_one()) {
      // ================================================================= TMA producer of the dY tiles (B operand)
      int bs = 0;
      uint32_t bph = 0;
      for (int vt = blockIdx.x; vt < p.num_vtiles; vt += gridDim.x) {
        int n, z0, y0, g;
        decode(vt, n, z0, y0, g);
        mbar_wait(bar_bempty + 8 * bs, bph ^ 1);
        const uint32_t b_dst = smem0 + bs * p.b_bytes;
        mbar_expect_tx(bar_bfull + 8 * bs, p.b_bytes);
        for (uint32_t i = 0; i < p.b_boxes; ++i)
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
              ::"r"(b_dst + i * 128u * 128u), "l"((uint64_t)&tmy64), "r"(bar_bfull + 8 * bs), "r"(4 * g * p.cout + (int)i * 64),
                "r"(y0), "r"(z0), "r"(n)
              : "memory");
        if (++bs == p.b_stages) { bs = 0; bph ^= 1; }
      }
    }
  } else if (warp >= 6) {
    if (elect_one()) {
      // ================================================================= TMA producers of the input atoms (A operand)
      // one thread sustains one bulk-tensor copy per ~450 cycles (tools/pipe_rates.py): warp 6 + j fetches atom j of every
      // M-block; all four arrive on the block's full barrier (a producer without an atom in the last block just arrives)
      const int j = warp - 6;
      const int aps = p.atoms_per_slab, kh = p.kh, a_stages = p.a_stages, q_total = p.q_total;
      int at0, dy0, dz0;
      {
        const int q = mb0 * 4 + j;
        const int slab = q / aps;
        at0 = q - slab * aps;
        dy0 = slab % kh; dz0 = slab / kh;
      }
      int as = 0;
      uint32_t aph = 0;
      for (int vt = blockIdx.x; vt < p.num_vtiles; vt += gridDim.x) {
        int n, z0, y0, g;
        decode(vt, n, z0, y0, g);
        const int e0 = (4 * g - pw) * p.cin;
        int at = at0, dy = dy0, dz = dz0;
        for (int mb = mb0; mb < mb1; ++mb) {
          mbar_wait(bar_aempty + 8 * as, aph ^ 1);
          const uint32_t fa = bar_afull + 8 * as;
          if (mb * 4 + j < q_total) {
            mbar_expect_tx(fa, kXAtomBytes);
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                ::"r"(smem0 + p.a_off + as * kXBlockBytes + j * kXAtomBytes), "l"((uint64_t)&tmx32), "r"(fa), "r"(e0 + at * 32),
                  "r"(y0 + dy - ph), "r"(z0 + dz - pd), "r"(n)
                : "memory");
          } else {
            mbar_arrive(fa);
          }
          at += 4;
          while (at >= aps) { at -= aps; if (++dy == kh) { dy = 0; ++dz; } }
          if (++as == a_stages) { as = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ================================================================= MMA issuer
      // A: M-major SWIZZLE_64B atoms (32 el), 16 K rows = 1024 B; B: N-major SWIZZLE_128B atoms (64 el), 16 rows = 2048 B
      const uint64_t tmpl_a = make_smem_desc(0, kXAtomBytes, 512u, kSwizzle64);
      const uint64_t tmpl_b = make_smem_desc(0, 128u * 128u, 1024u, kSwizzle128);
      const uint32_t a0 = in_reg((smem0 + p.a_off) >> 4), b0 = in_reg(smem0 >> 4), b16 = in_reg(p.b_bytes >> 4);
      const uint32_t idesc = in_reg(p.idesc), nt = in_reg((uint32_t)p.nt);
      const int a_stages = in_reg(p.a_stages), b_stages = in_reg(p.b_stages), num_vtiles = in_reg(p.num_vtiles);
      const int stride = in_reg((int)gridDim.x), nmb = in_reg(mb1 - mb0);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0, acc = 0;
      bool a_ready = mbar_try_wait(bar_afull, 0);
      for (int vt = blockIdx.x; vt < num_vtiles; vt += stride) {
        mbar_wait(bar_bfull + 8 * bs, bph);
        const uint64_t bd = tmpl_b + (uint64_t)(b0 + (uint32_t)bs * b16);
        uint32_t d_tmem = tmem;
        for (int mb = 0; mb < nmb; ++mb, d_tmem += nt) {
          if (!a_ready) mbar_wait(bar_afull + 8 * as, aph);
          tc_fence_after();
          const uint64_t ad = tmpl_a + (uint64_t)(a0 + (uint32_t)as * (kXBlockBytes >> 4));
          const uint32_t cur = bar_aempty + 8 * as;
          if (++as == a_stages) { as = 0; aph ^= 1; }
          a_ready = mbar_try_wait(bar_afull + 8 * as, aph);            // latency overlaps the issue below
          umma_ksteps_strided<8>(d_tmem, ad, bd, 1024u >> 4, 2048u >> 4, idesc, acc);
          umma_commit(cur);
        }
        umma_commit(bar_bempty + 8 * bs);
        if (++bs == b_stages) { bs = 0; bph ^= 1; }
        acc = 1;
      }
      umma_commit(bar_done);
    }
  } else {
    // =================================================================== epilogue: fold Toeplitz diagonals into dw
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    const int taps = p.kd * p.kh * p.kw;
    mbar_wait(bar_done, 0);
    tc_fence_after();
    for (int mb = mb0; mb < mb1; ++mb) {
      const int q = mb * 4 + row / 32;                  // atom index
      const bool valid = q < p.q_total;
      const int slab = valid ? q / p.atoms_per_slab : 0;
      const int e = valid ? (q - slab * p.atoms_per_slab) * 32 + (row & 31) : 0;      // element inside the window
      const int xi = e / p.cin, ci = e - xi * p.cin;
      const uint32_t taddr = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)((mb - mb0) * p.nt);
      int j = 0, co = 0;
      for (int c0 = 0; c0 < p.nt; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + c0, r);
        tmem_ld_wait();
        const int kx = xi - j;
        if (valid && kx >= 0 && kx < p.kw) {
          const int tap = slab * p.kw + kx;
          float* dst = dw + ((int64_t)co * taps + tap) * p.cin + ci;
#pragma unroll
          for (int i = 0; i < 16; ++i) atomicAdd(dst + (int64_t)i * taps * p.cin, __uint_as_float(r[i]));
        }
        co += 16;
        if (co == p.cout) { co = 0; ++j; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------ x-folded wgrad, z-slab variant
// Same GEMM as conv_wgrad_xfold_kernel, but the input window is fetched as SLAB atoms: for one dy and one 32-element atom
// of the window, (8 + kd - 1) z-lines x 16 y-rows x 64 B.  The kd taps in z read the same slab atom with the descriptor
// start moved by whole 16-row lines (1 KB = two 64B-swizzle atoms, so the TMA-written swizzle stays valid).  Four slab
// atoms (consecutive in (dy, atom) order, LBO = slab-atom size) form the M = 128 operand of kd M-blocks, one per dz.
// A bytes per voxel tile drop ~2.4x; the old kernel ran at ~65 % of L2 bandwidth (profiles/README.md).
struct WgradSParams {
  int n, d, h, w, cin, cout;
  int kd, kh, kw;
  int zl;                               // z-lines per slab atom = 8 + kd - 1
  int groups_x, tiles_h, tiles_d, num_vtiles;
  int aps;                              // 32-element atoms per window row = kxp / 32 (xfold_geom)
  int xoff;                             // extra voxels at the window start (image-fed layers)
  int sa_total;                         // slab atoms = kh * aps
  int grp_total, gpc;                   // groups of 4 slab atoms; groups per CTA (blockIdx.y)
  int a_stages, b_stages;
  int nt;                               // 4 * cout
  uint32_t sa_bytes;                    // one slab atom: zl * 16 rows * 64 B
  uint32_t b_boxes, b_bytes, a_off;
  uint32_t idesc, tmem_cols;
  long long* dbg;                       // diagnostics (B200_DBG): per-CTA cycle counters of the MMA lane, or nullptr
};

template <typename T>
__global__ void __launch_bounds__(320, 1)
conv_wgrad_xslab_kernel(const __grid_constant__ CUtensorMap tmx32, const __grid_constant__ CUtensorMap tmy64,
                        float* __restrict__ dw, const WgradSParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_bar[2 * kMaxAStages + 2 * kMaxBStages + 1];
  __shared__ uint32_t s_tmem;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_afull = smem_u32(&s_bar[0]);
  const uint32_t bar_aempty = smem_u32(&s_bar[kMaxAStages]);
  const uint32_t bar_bfull = smem_u32(&s_bar[2 * kMaxAStages]);
  const uint32_t bar_bempty = smem_u32(&s_bar[2 * kMaxAStages + kMaxBStages]);
  const uint32_t bar_done = smem_u32(&s_bar[2 * kMaxAStages + 2 * kMaxBStages]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(bar_afull + 8 * s, 4); mbar_init(bar_aempty + 8 * s, 1); }   // 4 atom producers
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(bar_bfull + 8 * s, 1); mbar_init(bar_bempty + 8 * s, 1); }
    mbar_init(bar_done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmx32);
    tma_prefetch_desc(&tmy64);
  }
  if (warp == 1) tmem_alloc(smem_u32(&s_tmem), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  const int pd = p.kd / 2, ph = p.kh / 2, pw = p.kw / 2;
  const int g0 = blockIdx.y * p.gpc;
  const int g1 = min(g0 + p.gpc, p.grp_total);
  const uint32_t grp_bytes = 4u * p.sa_bytes;

  auto decode = [&](int vt, int& n, int& z0, int& y0, int& g) {
    int t = vt;
    g = t % p.groups_x; t /= p.groups_x;
    y0 = (t % p.tiles_h) * 16; t /= p.tiles_h;
    z0 = (t % p.tiles_d) * 8; t /= p.tiles_d;
    n = t;
  };

  if (warp == 0) {
    if (elect_one()) {
      // ================================================================= TMA producer of the dY tiles (B operand)
      int bs = 0;
      uint32_t bph = 0;
      for (int vt = blockIdx.x; vt < p.num_vtiles; vt += gridDim.x) {
        int n, z0, y0, g;
        decode(vt, n, z0, y0, g);
        mbar_wait(bar_bempty + 8 * bs, bph ^ 1);
        const uint32_t b_dst = smem0 + bs * p.b_bytes;
        mbar_expect_tx(bar_bfull + 8 * bs, p.b_bytes);
        for (uint32_t i = 0; i < p.b_boxes; ++i)
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
              ::"r"(b_dst + i * 128u * 128u), "l"((uint64_t)&tmy64), "r"(bar_bfull + 8 * bs), "r"(4 * g * p.cout + (int)i * 64),
                "r"(y0), "r"(z0), "r"(n)
              : "memory");
        if (++bs == p.b_stages) { bs = 0; bph ^= 1; }
        // dY is streamed from DRAM: pull the tile this CTA needs a few iterations from now into L2
        const int vt2 = vt + 4 * (int)gridDim.x;
        if (vt2 < p.num_vtiles) {
          int n2, z2, y2, g2;
          decode(vt2, n2, z2, y2, g2);
          for (uint32_t i = 0; i < p.b_boxes; ++i)
            asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];"
                         ::"l"((uint64_t)&tmy64), "r"(4 * g2 * p.cout + (int)i * 64), "r"(y2), "r"(z2), "r"(n2)
                         : "memory");
        }
      }
    }
  } else if (warp >= 6) {
    if (elect_one()) {
      // ================================================================= TMA producers of the slab atoms (A operand)
      // warp 6 + j fetches slab atom j of every group (one thread sustains one bulk-tensor copy per ~450 cycles)
      const int j = warp - 6;
      const int aps = p.aps, a_stages = p.a_stages, sa_total = p.sa_total;
      int at0, dy0;
      {
        const int s = g0 * 4 + j;
        dy0 = s / aps;
        at0 = s - dy0 * aps;
      }
      int as = 0;
      uint32_t aph = 0;
      for (int vt = blockIdx.x; vt < p.num_vtiles; vt += gridDim.x) {
        int n, z0, y0, g;
        decode(vt, n, z0, y0, g);
        const int e0 = (4 * g - pw - p.xoff) * p.cin;
        int at = at0, dy = dy0;
        for (int gi = g0; gi < g1; ++gi) {
          mbar_wait(bar_aempty + 8 * as, aph ^ 1);
          const uint32_t fa = bar_afull + 8 * as;
          if (gi * 4 + j < sa_total) {
            mbar_expect_tx(fa, p.sa_bytes);
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                ::"r"(smem0 + p.a_off + as * grp_bytes + j * p.sa_bytes), "l"((uint64_t)&tmx32), "r"(fa), "r"(e0 + at * 32),
                  "r"(y0 + dy - ph), "r"(z0 - pd), "r"(n)
                : "memory");
          } else {
            mbar_arrive(fa);
          }
          at += 4;
          while (at >= aps) { at -= aps; ++dy; }
          if (++as == a_stages) { as = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ================================================================= MMA issuer
      // A: M-major SWIZZLE_64B atoms (32 el), LBO = slab atom, one K step = 16 rows = 1 KB; tap dz starts dz lines in.
      // B: N-major SWIZZLE_128B atoms (64 el), one K step = 16 rows = 2 KB
      const uint64_t tmpl_a = make_smem_desc(0, p.sa_bytes, 512u, kSwizzle64);
      const uint64_t tmpl_b = make_smem_desc(0, 128u * 128u, 1024u, kSwizzle128);
      const uint32_t ta_lo = in_reg((uint32_t)tmpl_a), ta_hi = in_reg((uint32_t)(tmpl_a >> 32));
      const uint32_t tb_lo = in_reg((uint32_t)tmpl_b), tb_hi = in_reg((uint32_t)(tmpl_b >> 32));
      const uint32_t a0 = in_reg((smem0 + p.a_off) >> 4), b0 = in_reg(smem0 >> 4), b16 = in_reg(p.b_bytes >> 4);
      const uint32_t grp16 = in_reg(grp_bytes >> 4), idesc = in_reg(p.idesc), nt = in_reg((uint32_t)p.nt);
      const int a_stages = in_reg(p.a_stages), b_stages = in_reg(p.b_stages), num_vtiles = in_reg(p.num_vtiles);
      const int stride = in_reg((int)gridDim.x), ngrp = in_reg(g1 - g0), kd = in_reg(p.kd);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0, acc = 0;
      bool a_ready = mbar_try_wait(bar_afull, 0);
      const bool dbg = p.dbg != nullptr;
      long long c_a = 0, c_b = 0;
      const long long c_start = clock64();
      for (int vt = blockIdx.x; vt < num_vtiles; vt += stride) {
        {
          const long long c0 = dbg ? clock64() : 0;
          mbar_wait(bar_bfull + 8 * bs, bph);
          if (dbg) c_b += clock64() - c0;
        }
        const uint32_t bd_lo = tb_lo + b0 + (uint32_t)bs * b16;
        uint32_t d_tmem = tmem;
        for (int gi = 0; gi < ngrp; ++gi) {
          if (!a_ready) {
            const long long c0 = dbg ? clock64() : 0;
            mbar_wait(bar_afull + 8 * as, aph);
            if (dbg) c_a += clock64() - c0;
          }
          tc_fence_after();
          uint32_t ad_lo = ta_lo + a0 + (uint32_t)as * grp16;
          const uint32_t cur = bar_aempty + 8 * as;
          if (++as == a_stages) { as = 0; aph ^= 1; }
          a_ready = mbar_try_wait(bar_afull + 8 * as, aph);            // latency overlaps the issue below
          for (int dz = 0; dz < kd; ++dz, ad_lo += 1024u >> 4, d_tmem += nt)
            umma_ksteps_split<8>(d_tmem, ad_lo, ta_hi, bd_lo, tb_hi, 1024u >> 4, 2048u >> 4, idesc, acc);
          umma_commit(cur);
        }
        umma_commit(bar_bempty + 8 * bs);
        if (++bs == b_stages) { bs = 0; bph ^= 1; }
        acc = 1;
      }
      umma_commit(bar_done);
      if (dbg) {
        long long* o = p.dbg + (blockIdx.y * gridDim.x + blockIdx.x) * 4;
        o[0] = clock64() - c_start; o[1] = c_a; o[2] = c_b;
      }
    }
  } else {
    // =================================================================== epilogue: fold Toeplitz diagonals into dw
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    const int taps = p.kd * p.kh * p.kw;
    mbar_wait(bar_done, 0);
    tc_fence_after();
    for (int gi = g0; gi < g1; ++gi) {
      const int s = gi * 4 + row / 32;                  // slab atom of this row
      const bool valid = s < p.sa_total;
      const int dyy = valid ? s / p.aps : 0;
      const int e = valid ? (s - dyy * p.aps) * 32 + (row & 31) : 0;      // element inside the window
      const int xi = e / p.cin, ci = e - xi * p.cin;
      for (int dz = 0; dz < p.kd; ++dz) {
        const uint32_t taddr = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(((gi - g0) * p.kd + dz) * p.nt);
        int j = 0, co = 0;
        for (int c0 = 0; c0 < p.nt; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(taddr + c0, r);
          tmem_ld_wait();
          const int kx = xi - p.xoff - j;
          if (valid && kx >= 0 && kx < p.kw) {
            const int tap = (dz * p.kh + dyy) * p.kw + kx;
            float* dst = dw + ((int64_t)co * taps + tap) * p.cin + ci;
#pragma unroll
            for (int i = 0; i < 16; ++i) atomicAdd(dst + (int64_t)i * taps * p.cin, __uint_as_float(r[i]));
          }
          co += 16;
          if (co == p.cout) { co = 0; ++j; }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------- wgrad, cp.async-fed
// Same decomposition as conv_wgrad_umma_kernel; operands are written by four loader warps (thread r <-> voxel row r)
// with cp.async into the 32-byte-swizzled chunk tiles (A) and the 32/64/128-byte-swizzled dY tile (B).
// Warp roles (288 threads): warp 0 = MMA issuer + TMEM owner, warps 1-4 = epilogue (once, at the end), 5-8 = loaders.
struct WgradParams2 {
  WgradParams base;
  int64_t xsw, xsh, xsd, xsn;
  int64_t ysw, ysh, ysd, ysn;
};

template <typename T>
__global__ void __launch_bounds__(288, 1)
conv_wgrad_umma2_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw, const WgradParams2 pp) {
  const WgradParams& p = pp.base;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_bar[2 * kMaxAStages + 2 * kMaxBStages + 1];
  __shared__ uint32_t s_tmem;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_afull = smem_u32(&s_bar[0]);
  const uint32_t bar_aempty = smem_u32(&s_bar[kMaxAStages]);
  const uint32_t bar_bfull = smem_u32(&s_bar[2 * kMaxAStages]);
  const uint32_t bar_bempty = smem_u32(&s_bar[2 * kMaxAStages + kMaxBStages]);
  const uint32_t bar_done = smem_u32(&s_bar[2 * kMaxAStages + 2 * kMaxBStages]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(bar_afull + 8 * s, 128); mbar_init(bar_aempty + 8 * s, 1); }
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(bar_bfull + 8 * s, 128); mbar_init(bar_bempty + 8 * s, 1); }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  const int pd = p.kd / 2, ph = p.kh / 2, pw = p.kw / 2;
  const int mb0 = blockIdx.y * p.g;
  const int mb1 = min(mb0 + p.g, p.mb_total);

  auto decode = [&](int vt, int& n, int& z0, int& y0, int& x0) {
    int t = vt;
    x0 = (t % p.tiles_w) * p.bw; t /= p.tiles_w;
    y0 = (t % p.tiles_h) * p.bh; t /= p.tiles_h;
    z0 = (t % p.tiles_d) * p.bd; t /= p.tiles_d;
    n = t;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ================================================================= MMA issuer
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      bool first = true;
      for (int vt = blockIdx.x; vt < p.num_vtiles; vt += gridDim.x) {
        mbar_wait(bar_bfull + 8 * bs, bph);
        fence_proxy_async();
        tc_fence_after();
        const uint32_t b_base = smem0 + bs * p.b_bytes;
        for (int mb = mb0; mb < mb1; ++mb) {
          mbar_wait(bar_afull + 8 * as, aph);
          fence_proxy_async();
          tc_fence_after();
          const uint32_t a_base = smem0 + p.a_off + as * kBlockBytes;
          const uint32_t d_tmem = tmem + (uint32_t)((mb - mb0) * p.cout);
#pragma unroll 1
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t ad = make_smem_desc(a_base + ks * 512u, kChunkBytes, 256u, kSwizzle32);
            const uint64_t bd = make_smem_desc(b_base + ks * p.b_kstep, p.b_lbo, p.b_sbo, p.b_layout);
            umma_f16(d_tmem, ad, bd, p.idesc, (first && ks == 0) ? 0u : 1u);
          }
          umma_commit(bar_aempty + 8 * as);
          if (++as == p.a_stages) { as = 0; aph ^= 1; }
        }
        umma_commit(bar_bempty + 8 * bs);
        if (++bs == p.b_stages) { bs = 0; bph ^= 1; }
        first = false;
      }
      umma_commit(bar_done);
    }
  } else if (warp <= 4) {
    // =================================================================== epilogue: TMEM -> fp32 atomics into dw
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    const int taps = p.kd * p.kh * p.kw;
    mbar_wait(bar_done, 0);
    tc_fence_after();
    for (int mb = mb0; mb < mb1; ++mb) {
      const int q = mb * 8 + row / 16;
      const bool valid = q < p.q_total;
      const int tap = valid ? q / p.chunks16 : 0;
      const int ci = valid ? (q - tap * p.chunks16) * 16 + (row & 15) : 0;
      const uint32_t taddr = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)((mb - mb0) * p.cout);
      for (int j0 = 0; j0 < p.cout; j0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + j0, r);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; ++j) atomicAdd(dw + ((int64_t)(j0 + j) * taps + tap) * p.cin + ci, __uint_as_float(r[j]));
        }
      }
    }
  } else {
    // =================================================================== loaders
    const int r = threadIdx.x - 160;
    const int lx = r % p.bw, ly = (r / p.bw) % p.bh, lz = r / (p.bw * p.bh);
    const uint32_t rp = p.b_box_c * 2;                 // dY row pitch inside one box tile
    const int bcpr = (int)(p.b_box_c / 8);             // 16-byte chunks per dY box row
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    for (int vt = blockIdx.x; vt < p.num_vtiles; vt += gridDim.x) {
      int n, z0, y0, x0;
      decode(vt, n, z0, y0, x0);
      const int gz = z0 + lz, gy = y0 + ly, gx = x0 + lx;
      const bool inside = gz < p.d && gy < p.h && gx < p.w;
      // ---- dY tile (B operand)
      mbar_wait(bar_bempty + 8 * bs, bph ^ 1);
      {
        const uint32_t b_dst = smem0 + bs * p.b_bytes;
        const T* yrow = inside ? dy + (int64_t)n * pp.ysn + (int64_t)gz * pp.ysd + (int64_t)gy * pp.ysh + (int64_t)gx * pp.ysw : dy;
        for (uint32_t i = 0; i < p.b_boxes; ++i) {
          const uint32_t rowaddr = b_dst + i * 128u * rp + (uint32_t)r * rp;
          for (int j = 0; j < bcpr; ++j) {
            const uint32_t sj = rp == 128 ? ((uint32_t)j ^ ((uint32_t)r & 7u))
                                          : (rp == 64 ? ((uint32_t)j ^ (((uint32_t)r >> 1) & 3u)) : ((uint32_t)j ^ (((uint32_t)r >> 2) & 1u)));
            cp_async16(rowaddr + (sj << 4), yrow + i * p.b_box_c + j * 8, inside ? 16u : 0u);
          }
        }
        cp_async_arrive_noinc(bar_bfull + 8 * bs);
        if (++bs == p.b_stages) { bs = 0; bph ^= 1; }
      }
      // ---- activation chunk tiles (A operand), 8 per M-block
      const T* xrow = x + (int64_t)n * pp.xsn + (int64_t)gz * pp.xsd + (int64_t)gy * pp.xsh + (int64_t)gx * pp.xsw;
      int cc, dx, dyy, dz;
      {
        const int q = mb0 * 8;
        const int tap = q / p.chunks16;
        cc = q - tap * p.chunks16;
        dx = tap % p.kw; dyy = (tap / p.kw) % p.kh; dz = tap / (p.kw * p.kh);
      }
      const uint32_t sw = ((uint32_t)r >> 2) & 1u;
      for (int mb = mb0; mb < mb1; ++mb) {
        mbar_wait(bar_aempty + 8 * as, aph ^ 1);
        const uint32_t a_dst = smem0 + p.a_off + as * kBlockBytes;
        const int nq = min(8, p.q_total - mb * 8);
        uint32_t rowaddr = a_dst + (uint32_t)r * 32u;
        for (int j = 0; j < nq; ++j, rowaddr += kChunkBytes) {
          const int iz = gz + dz - pd, iy = gy + dyy - ph, ix = gx + dx - pw;
          const bool ok = iz >= 0 && iz < p.d && iy >= 0 && iy < p.h && ix >= 0 && ix < p.w;
          const T* src = ok ? xrow + (int64_t)(dz - pd) * pp.xsd + (int64_t)(dyy - ph) * pp.xsh + (int64_t)(dx - pw) * pp.xsw + cc * 16 : x;
          cp_async16(rowaddr + (sw << 4), src, ok ? 16u : 0u);
          cp_async16(rowaddr + ((sw ^ 1u) << 4), src + 8, ok ? 16u : 0u);
          if (++cc == p.chunks16) { cc = 0; if (++dx == p.kw) { dx = 0; if (++dyy == p.kh) { dyy = 0; ++dz; } } }
        }
        cp_async_arrive_noinc(bar_afull + 8 * as);
        if (++as == p.a_stages) { as = 0; aph ^= 1; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, p.tmem_cols);
  }
}

bool conv_wgrad_xfold_ok(const ActView& x, const ActView& dy, int kd, int kh, int kw) {
  if (x.dtype != B200_BF16 && x.dtype != B200_F16) return false;
  if ((kw != 3 && kw != 1) || kd > 5 || kh > 5) return false;
  if (x.c > 96) return false;
  if (x.c % 16 != 0) {
    int xoff, kxp;
    if (!xfold_geom(x.c, kw, &xoff, &kxp)) return false;
    if (kd != 3 || x.d < 8 || x.h < 16 || kd * 4 * dy.c > 512) return false;      // z-slab kernel only
  } else if (((3 + kw) * x.c) % 32 != 0) {
    return false;
  }
  if (!(dy.c == 16 || dy.c == 32 || dy.c == 64)) return false;
  if (x.sw != x.c || x.sh != (int64_t)x.w * x.c) return false;
  if (dy.sw != dy.c || dy.sh != (int64_t)dy.w * dy.c) return false;
  if (x.w % 4 != 0 || x.w < 8) return false;
  if (!aligned16(x.data) || !aligned16(dy.data)) return false;
  if ((int64_t)x.d * x.h < 128) return false;
  return encode_tiled_fn() != nullptr;
}

static int conv_wgrad_xslab_v(const ActView& x, const ActView& dy, float* dw, int kd, int kh, int kw, cudaStream_t st) {
  WgradSParams p{};
  p.n = x.n; p.d = x.d; p.h = x.h; p.w = x.w; p.cin = x.c; p.cout = dy.c;
  p.kd = kd; p.kh = kh; p.kw = kw;
  p.zl = 8 + kd - 1;
  p.groups_x = x.w / 4;
  p.tiles_h = (int)ceil_div(x.h, 16); p.tiles_d = (int)ceil_div(x.d, 8);
  p.num_vtiles = x.n * p.tiles_d * p.tiles_h * p.groups_x;
  int kxp = 0;
  B200_CHECK_ARG(xfold_geom(x.c, kw, &p.xoff, &kxp), "conv_wgrad(xslab): unsupported Cin");
  p.aps = kxp / 32;
  p.sa_total = kh * p.aps;
  p.grp_total = (int)ceil_div(p.sa_total, 4);
  p.nt = 4 * dy.c;
  // groups per CTA: TMEM holds gpc * kd accumulators of nt columns; among the feasible values take the one that
  // minimises (M-blocks per CTA) / (CTAs along the voxel-tile axis)
  const int gpc_max = 512 / (kd * p.nt);
  B200_CHECK_ARG(gpc_max >= 1, "conv_wgrad(xslab): accumulators do not fit in TMEM");
  double best = 1e30;
  for (int gpc = 1; gpc <= gpc_max && gpc <= p.grp_total; ++gpc) {
    const int gy = (int)ceil_div(p.grp_total, gpc);
    int vs = sm_count() / gy;
    if (vs < 1) vs = 1;
    const double cost = (double)gpc / vs;
    if (cost < best - 1e-12) { best = cost; p.gpc = gpc; }
  }
  const int gy = (int)ceil_div(p.grp_total, p.gpc);
  p.sa_bytes = (uint32_t)p.zl * 16u * 64u;
  p.b_boxes = (uint32_t)p.nt / 64u;
  p.b_bytes = 128u * (uint32_t)p.nt * 2u;
  // the dY tiles come from DRAM (each is read once per CTA row): give that ring the depth, keep 3 slab groups in flight
  int a_st = 3;
  int b_st = (int)((200u * 1024u - (uint32_t)a_st * 4u * p.sa_bytes) / p.b_bytes);
  if (b_st < 2) { a_st = 2; b_st = (int)((200u * 1024u - (uint32_t)a_st * 4u * p.sa_bytes) / p.b_bytes); }
  if (b_st > kMaxBStages) b_st = kMaxBStages;
  B200_CHECK_ARG(b_st >= 2, "conv_wgrad(xslab): tiles do not fit in shared memory");
  p.b_stages = b_st;
  p.a_off = p.b_stages * p.b_bytes;
  a_st = (int)((200u * 1024u - p.a_off) / (4u * p.sa_bytes));
  if (a_st > kMaxAStages) a_st = kMaxAStages;
  p.a_stages = a_st;
  p.idesc = make_idesc(x.dtype == B200_BF16, p.nt, 1, 1);
  uint32_t cols = 32;
  while (cols < (uint32_t)(p.gpc * kd * p.nt)) cols <<= 1;
  p.tmem_cols = cols;

  CUtensorMap tx, ty;
  const cuuint64_t xd[4] = {(cuuint64_t)x.w * x.c, (cuuint64_t)x.h, (cuuint64_t)x.d, (cuuint64_t)x.n};
  const cuuint64_t xs[3] = {(cuuint64_t)x.sh * 2, (cuuint64_t)x.sd * 2, (cuuint64_t)x.sn * 2};
  const cuuint32_t bx[4] = {32, 16, (cuuint32_t)p.zl, 1};
  int rc = make_tmap4(&tx, x.data, x.dtype, xd, xs, bx);
  if (rc) return rc;
  const cuuint64_t yd[4] = {(cuuint64_t)dy.w * dy.c, (cuuint64_t)dy.h, (cuuint64_t)dy.d, (cuuint64_t)dy.n};
  const cuuint64_t ys[3] = {(cuuint64_t)dy.sh * 2, (cuuint64_t)dy.sd * 2, (cuuint64_t)dy.sn * 2};
  const cuuint32_t by[4] = {64, 16, 8, 1};
  rc = make_tmap4(&ty, dy.data, dy.dtype, yd, ys, by);
  if (rc) return rc;

  int vsplit = sm_count() / gy;
  if (vsplit < 1) vsplit = 1;
  if (vsplit > p.num_vtiles) vsplit = p.num_vtiles;
  dim3 grid((unsigned)vsplit, (unsigned)gy);
  static long long* dbg_buf = nullptr;
  if (getenv("B200_DBG")) {
    if (!dbg_buf) B200_CUDA(cudaMalloc(&dbg_buf, sizeof(long long) * 4 * 1024));
    B200_CUDA(cudaMemsetAsync(dbg_buf, 0, sizeof(long long) * 4 * 1024, st));
    p.dbg = dbg_buf;
  }
  const size_t smem = (size_t)p.a_off + (size_t)p.a_stages * 4u * p.sa_bytes + 1024;
  if (x.dtype == B200_BF16) {
    auto kern = conv_wgrad_xslab_kernel<__nv_bfloat16>;
    B200_CUDA(raise_dyn_smem_cap(kern));
    kern<<<grid, 320, smem, st>>>(tx, ty, dw, p);
  } else {
    auto kern = conv_wgrad_xslab_kernel<__half>;
    B200_CUDA(raise_dyn_smem_cap(kern));
    kern<<<grid, 320, smem, st>>>(tx, ty, dw, p);
  }
  B200_LAUNCH_CHECK();
  if (p.dbg) {
    long long h[4 * 1024];
    B200_CUDA(cudaMemcpyAsync(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    double tot = 0, ca = 0, cb = 0;
    const int nct = vsplit * gy;
    for (int b = 0; b < nct && b < 1024; ++b) { tot += h[b * 4]; ca += h[b * 4 + 1]; cb += h[b * 4 + 2]; }
    const double per = (double)p.num_vtiles * gy;      // (voxel tile, CTA row) pairs
    printf("wgrad xslab dbg: grid %dx%d gpc %d stages a%d b%d | per voxel tile and CTA: MMA-lane loop %.0f cycles, wait slab atoms %.0f, "
           "wait dY %.0f\n", vsplit, gy, p.gpc, p.a_stages, p.b_stages, tot / per, ca / per, cb / per);
    fflush(stdout);
  }
  return B200_OK;
}

int conv_wgrad_xfold_v(const ActView& x, const ActView& dy, float* dw, int kd, int kh, int kw, cudaStream_t st) {
  B200_CHECK_ARG(conv_wgrad_xfold_ok(x, dy, kd, kh, kw), "conv_wgrad(xfold): unsupported operands");
  if (xslab_mode() && kd == 3 && x.d >= 8 && x.h >= 16 && kd * 4 * dy.c <= 512)
    return conv_wgrad_xslab_v(x, dy, dw, kd, kh, kw, st);
  WgradXParams p{};
  p.n = x.n; p.d = x.d; p.h = x.h; p.w = x.w; p.cin = x.c; p.cout = dy.c;
  p.kd = kd; p.kh = kh; p.kw = kw; p.win = 3 + kw;
  {
    static const int cand[][2] = {{16, 8}, {8, 16}, {32, 4}, {4, 32}, {64, 2}, {2, 64}, {128, 1}, {1, 128}};
    double best = 1e30;
    for (auto& c : cand) {
      double cover = (double)ceil_div(x.h, c[0]) * c[0] * ceil_div(x.d, c[1]) * c[1];
      if (cover < best - 0.5) { best = cover; p.bh = c[0]; p.bd = c[1]; }
    }
  }
  p.groups_x = x.w / 4;
  p.tiles_h = (int)ceil_div(x.h, p.bh); p.tiles_d = (int)ceil_div(x.d, p.bd);
  p.num_vtiles = x.n * p.tiles_d * p.tiles_h * p.groups_x;
  p.atoms_per_slab = p.win * x.c / 32;
  p.q_total = kd * kh * p.atoms_per_slab;
  p.mb_total = (int)ceil_div(p.q_total, 4);
  p.nt = 4 * dy.c;
  p.g = 512 / p.nt;
  if (p.g > p.mb_total) p.g = p.mb_total;
  const int groups = (int)ceil_div(p.mb_total, p.g);
  p.b_boxes = (uint32_t)p.nt / 64u;
  p.b_bytes = 128u * (uint32_t)p.nt * 2u;
  p.b_stages = 2;
  p.a_off = p.b_stages * p.b_bytes;
  int a_st = (int)((200u * 1024u - p.a_off) / kXBlockBytes);
  if (a_st > 4) a_st = 4;
  B200_CHECK_ARG(a_st >= 2, "conv_wgrad(xfold): tiles do not fit in shared memory");
  p.a_stages = a_st;
  p.idesc = make_idesc(x.dtype == B200_BF16, p.nt, 1, 1);
  uint32_t cols = 32;
  while (cols < (uint32_t)(p.g * p.nt)) cols <<= 1;
  p.tmem_cols = cols;

  CUtensorMap tx, ty;
  const cuuint64_t xd[4] = {(cuuint64_t)x.w * x.c, (cuuint64_t)x.h, (cuuint64_t)x.d, (cuuint64_t)x.n};
  const cuuint64_t xs[3] = {(cuuint64_t)x.sh * 2, (cuuint64_t)x.sd * 2, (cuuint64_t)x.sn * 2};
  const cuuint32_t bx[4] = {32, (cuuint32_t)p.bh, (cuuint32_t)p.bd, 1};
  int rc = make_tmap4(&tx, x.data, x.dtype, xd, xs, bx);
  if (rc) return rc;
  const cuuint64_t yd[4] = {(cuuint64_t)dy.w * dy.c, (cuuint64_t)dy.h, (cuuint64_t)dy.d, (cuuint64_t)dy.n};
  const cuuint64_t ys[3] = {(cuuint64_t)dy.sh * 2, (cuuint64_t)dy.sd * 2, (cuuint64_t)dy.sn * 2};
  const cuuint32_t by[4] = {64, (cuuint32_t)p.bh, (cuuint32_t)p.bd, 1};
  rc = make_tmap4(&ty, dy.data, dy.dtype, yd, ys, by);
  if (rc) return rc;

  int vsplit = sm_count() / groups;
  if (vsplit < 1) vsplit = 1;
  if (vsplit > p.num_vtiles) vsplit = p.num_vtiles;
  dim3 grid((unsigned)vsplit, (unsigned)groups);
  const size_t smem = (size_t)p.a_off + (size_t)p.a_stages * kXBlockBytes + 1024;
  if (x.dtype == B200_BF16) {
    auto kern = conv_wgrad_xfold_kernel<__nv_bfloat16>;
    B200_CUDA(raise_dyn_smem_cap(kern));
    kern<<<grid, 320, smem, st>>>(tx, ty, dw, p);
  } else {
    auto kern = conv_wgrad_xfold_kernel<__half>;
    B200_CUDA(raise_dyn_smem_cap(kern));
    kern<<<grid, 320, smem, st>>>(tx, ty, dw, p);
  }
  B200_LAUNCH_CHECK();
  return B200_OK;
}

}  // namespace sm100

bool conv_wgrad_umma_supported(const b200_tensor* x, const b200_tensor* dy, int kd, int kh, int kw) {
  if (x->dtype != B200_BF16 && x->dtype != B200_F16) return false;
  if (x->c % 16 != 0 && x->c <= 8 && dy->ld % 8 == 0)
    return sm100::conv_wgrad_xfold_ok(sm100::view_of(x), sm100::view_of(dy), kd, kh, kw);   // image-fed layer: x-folded slab kernel
  if (x->c % 16 != 0 || x->ld % 8 != 0 || dy->ld % 8 != 0) return false;
  const int co = dy->c;
  if (!(co == 16 || co == 32 || co == 64 || co == 128 || co == 256)) return false;
  if (!sm100::aligned16(x->data) || !sm100::aligned16(dy->data)) return false;
  if (kd * kh * kw > 125) return false;
  if ((int64_t)x->n * x->d * x->h * x->w < 128) return false;
  return sm100::encode_tiled_fn() != nullptr;
}

static int wgrad_xfold_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("B200_XFOLD");
    mode = (e && strcmp(e, "0") == 0) ? 0 : 1;
  }
  return mode;
}

int conv_wgrad_umma(const b200_tensor* x, const b200_tensor* dy, float* dw, float* dbias, int kd, int kh, int kw,
                    cudaStream_t st) {
  const sm100::ActView xv = sm100::view_of(x), dv = sm100::view_of(dy);
  int rc = (wgrad_xfold_mode() && sm100::conv_wgrad_xfold_ok(xv, dv, kd, kh, kw))
               ? sm100::conv_wgrad_xfold_v(xv, dv, dw, kd, kh, kw, st)
               : sm100::conv_wgrad_umma_v(xv, dv, dw, kd, kh, kw, st);
  if (rc) return rc;
  if (dbias) return conv_bias_grad(dy, dbias, st);
  return B200_OK;
}

namespace sm100 {
static int conv_wgrad_umma_impl(const ActView& xv, const ActView& dyv, float* dw, int kd, int kh, int kw, cudaStream_t st,
                                const ActView* phase_views, int nphases);

int conv_wgrad_umma_v(const ActView& xv, const ActView& dyv, float* dw, int kd, int kh, int kw, cudaStream_t st) {
  return conv_wgrad_umma_impl(xv, dyv, dw, kd, kh, kw, st, nullptr, 0);
}

// dw[t][co][ci] = sum_v phase_t[v][co] * x[v][ci] for all phases: the GEMM N dimension is (phase, co), <= 256 columns per launch
int conv_wgrad_phases_v(const ActView& x, const ActView* phase_views, int nphases, float* dw, cudaStream_t st) {
  const int co = phase_views[0].c;
  int per = 256 / co;
  if (per < 1) per = 1;
  if (per > 8) per = 8;
  for (int t0 = 0; t0 < nphases; t0 += per) {
    const int nt = nphases - t0 < per ? nphases - t0 : per;
    int rc = conv_wgrad_umma_impl(x, phase_views[t0], dw + (size_t)t0 * co * x.c, 1, 1, 1, st, phase_views + t0, nt);
    if (rc) return rc;
  }
  return B200_OK;
}

static int conv_wgrad_umma_impl(const ActView& xv, const ActView& dyv, float* dw, int kd, int kh, int kw, cudaStream_t st,
                                const ActView* phase_views, int nphases) {
  const ActView* x = &xv;
  const ActView* dy = &dyv;
  WgradParams p{};
  p.n = x->n; p.d = x->d; p.h = x->h; p.w = x->w; p.cin = x->c; p.cout = dy->c * (nphases ? nphases : 1);
  p.kd = kd; p.kh = kh; p.kw = kw;
  pick_tile(x->d, x->h, x->w, &p.bd, &p.bh, &p.bw);
  p.tiles_d = (int)ceil_div(x->d, p.bd); p.tiles_h = (int)ceil_div(x->h, p.bh); p.tiles_w = (int)ceil_div(x->w, p.bw);
  p.num_vtiles = x->n * p.tiles_d * p.tiles_h * p.tiles_w;
  p.chunks16 = x->c / 16;
  p.q_total = kd * kh * kw * p.chunks16;
  p.mb_total = (int)ceil_div(p.q_total, 8);
  p.g = 512 / p.cout;
  if (p.g > p.mb_total) p.g = p.mb_total;
  const int groups = (int)ceil_div(p.mb_total, p.g);
  // B operand (dy tile): channels per TMA box = min(cout, 64); N-major atoms of that width
  p.b_box_c = dy->c < 64 ? (uint32_t)dy->c : 64u;
  p.b_boxes = (uint32_t)p.cout / p.b_box_c;
  p.bpp = nphases ? dy->c / (int)p.b_box_c : 0;
  const uint32_t rp = p.b_box_c * 2;                  // row pitch in bytes = swizzle width
  p.b_layout = rp == 128 ? kSwizzle128 : (rp == 64 ? kSwizzle64 : kSwizzle32);
  p.b_sbo = 8u * rp;
  p.b_lbo = 128u * rp;
  p.b_kstep = 16u * rp;
  p.b_bytes = 128u * (uint32_t)p.cout * 2u;
  p.b_stages = 2;
  p.a_off = p.b_stages * p.b_bytes;
  int a_st = (int)((200u * 1024u - p.a_off) / kBlockBytes);
  if (a_st > 4) a_st = 4;
  B200_CHECK_ARG(a_st >= 2, "conv_wgrad(umma): tiles do not fit in shared memory");
  p.a_stages = a_st;
  p.idesc = make_idesc(x->dtype == B200_BF16, p.cout, 1, 1);
  uint32_t cols = 32;
  while (cols < (uint32_t)(p.g * p.cout)) cols <<= 1;
  p.tmem_cols = cols;

  int vsplit = sm_count() / groups;
  if (vsplit < 1) vsplit = 1;
  if (vsplit > p.num_vtiles) vsplit = p.num_vtiles;
  dim3 grid((unsigned)vsplit, (unsigned)groups);
  const size_t smem = (size_t)p.a_off + (size_t)p.a_stages * kBlockBytes + 1024;
  if (loader_mode() == 1 && nphases == 0) {
    WgradParams2 pp{p, x->sw, x->sh, x->sd, x->sn, dy->sw, dy->sh, dy->sd, dy->sn};
    if (x->dtype == B200_BF16) {
      auto kern = conv_wgrad_umma2_kernel<__nv_bfloat16>;
      B200_CUDA(raise_dyn_smem_cap(kern));
      kern<<<grid, 288, smem, st>>>((const __nv_bfloat16*)x->data, (const __nv_bfloat16*)dy->data, dw, pp);
    } else {
      auto kern = conv_wgrad_umma2_kernel<__half>;
      B200_CUDA(raise_dyn_smem_cap(kern));
      kern<<<grid, 288, smem, st>>>((const __half*)x->data, (const __half*)dy->data, dw, pp);
    }
    B200_LAUNCH_CHECK();
    return B200_OK;
  }
  CUtensorMap tx, tdy;
  int rc = make_act_tmap(&tx, *x, 16, p.bw, p.bh, p.bd);
  if (rc) return rc;
  rc = make_act_tmap(&tdy, *dy, (int)p.b_box_c, p.bw, p.bh, p.bd);
  if (rc) return rc;
  PhaseMaps pm;
  memset(&pm, 0, sizeof(pm));
  for (int t = 0; t < nphases; ++t) {
    rc = make_act_tmap(&pm.m[t], phase_views[t], (int)p.b_box_c, p.bw, p.bh, p.bd);
    if (rc) return rc;
  }
  if (x->dtype == B200_BF16) {
    auto kern = conv_wgrad_umma_kernel<__nv_bfloat16>;
    B200_CUDA(raise_dyn_smem_cap(kern));
    kern<<<grid, 320, smem, st>>>(tx, tdy, pm, dw, p);
  } else {
    auto kern = conv_wgrad_umma_kernel<__half>;
    B200_CUDA(raise_dyn_smem_cap(kern));
    kern<<<grid, 320, smem, st>>>(tx, tdy, pm, dw, p);
  }
  B200_LAUNCH_CHECK();
  return B200_OK;
}
}  // namespace sm100

}  // namespace b200

// ============================================================================== transposed convolution (k == s)
// Each of the s^3 output phases is a pointwise convolution between the coarse tensor and a strided sub-lattice view
// of the fine tensor, so the three passes reuse the kernels above with ActView strides:
//   fprop : for every phase t   Y_t   = X . W_t^T + b         (conv_fprop_umma_v, k = 1, strided epilogue)
//   dgrad : dX = sum_t dY_t . W_t                             (k = 1, TMA reads the strided view, accumulate epilogue)
//   wgrad : dW_t = X^T . dY_t                                 (conv_wgrad_umma_v, k = 1)
namespace b200 {
namespace sm100 {

template <typename T>
__global__ void pack_convT_weight_kernel(const float* __restrict__ w, T* __restrict__ p, int cin, int cout, int taps, int for_dgrad) {
  const int64_t total = (int64_t)cin * cout * taps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int t = (int)(i % taps);
    int co = (int)((i / taps) % cout);
    int ci = (int)(i / ((int64_t)taps * cout));
    int64_t o = for_dgrad ? (((int64_t)t * cin + ci) * cout + co) : (((int64_t)t * cout + co) * cin + ci);
    p[o] = from_f<T>(w[i]);
  }
}

__global__ void unpack_convT_wgrad_kernel(const float* __restrict__ p, float* __restrict__ dw, int cin, int cout, int taps,
                                          int accumulate) {
  const int64_t total = (int64_t)cin * cout * taps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int t = (int)(i % taps);
    int co = (int)((i / taps) % cout);
    int ci = (int)(i / ((int64_t)taps * cout));
    float v = p[((int64_t)t * cout + co) * cin + ci];
    dw[i] = accumulate ? dw[i] + v : v;
  }
}

static bool convT_tc_ok(const b200_tensor* x, const b200_tensor* y, int sd, int sh, int sw) {
  if (x->dtype != B200_BF16 && x->dtype != B200_F16) return false;
  if (y->dtype != x->dtype) return false;
  if (x->c % 16 != 0 || y->c % 16 != 0 || x->ld % 8 != 0 || y->ld % 8 != 0) return false;
  const int co = y->c;
  if (!(co == 16 || co == 32 || co == 64 || co == 128 || co == 256)) return false;   // wgrad N tile
  if (x->c > 256 && x->c % 128 != 0) return false;                                   // dgrad N tiles: Cin itself or 128 wide
  if (!aligned16(x->data) || !aligned16(y->data)) return false;
  if (y->d != x->d * sd || y->h != x->h * sh || y->w != x->w * sw) return false;
  if ((int64_t)x->n * x->d * x->h * x->w < 128) return false;
  return encode_tiled_fn() != nullptr;
}

}  // namespace sm100
}  // namespace b200

using namespace b200;
using namespace b200::sm100;

B200_EXPORT int b200_convT_tc_supported(const b200_tensor* x, const b200_tensor* y, int32_t sd, int32_t sh, int32_t sw) {
  if (!x || !y) return 0;
  return convT_tc_ok(x, y, sd, sh, sw) ? 1 : 0;
}

B200_EXPORT int b200_pack_convT_weight(const float* w, void* packed, int32_t dtype, int32_t cin, int32_t cout, int32_t taps,
                                       int32_t for_dgrad, void* stream) {
  B200_CHECK_ARG(w && packed && cin > 0 && cout > 0 && taps > 0, "pack_convT_weight: bad args");
  int64_t total = (int64_t)cin * cout * taps;
  unsigned blocks = (unsigned)(ceil_div(total, 256) < 4096 ? ceil_div(total, 256) : 4096);
  if (dtype == B200_BF16)
    pack_convT_weight_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)packed, cin, cout, taps, for_dgrad);
  else if (dtype == B200_F16)
    pack_convT_weight_kernel<__half><<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (__half*)packed, cin, cout, taps, for_dgrad);
  else {
    set_error("pack_convT_weight: 16-bit dtypes only");
    return B200_ERR_ARG;
  }
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_unpack_convT_wgrad(const float* dw_packed, float* dw, int32_t cin, int32_t cout, int32_t taps,
                                        int32_t accumulate, void* stream) {
  B200_CHECK_ARG(dw_packed && dw, "unpack_convT_wgrad: null pointer");
  int64_t total = (int64_t)cin * cout * taps;
  unsigned blocks = (unsigned)(ceil_div(total, 256) < 4096 ? ceil_div(total, 256) : 4096);
  unpack_convT_wgrad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dw_packed, dw, cin, cout, taps, accumulate);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

B200_EXPORT int b200_convT_fprop_tc(const b200_tensor* x, const void* w_packed, const float* bias, const b200_tensor* y,
                                    int32_t sd, int32_t sh, int32_t sw, void* stream) {
  B200_CHECK_ARG(check_tensor(x, "convT_fprop_tc.x") && check_tensor(y, "convT_fprop_tc.y") && w_packed, "%s", b200_last_error());
  B200_CHECK_ARG(convT_tc_ok(x, y, sd, sh, sw), "convT_fprop_tc: unsupported operands (query b200_convT_tc_supported first)");
  // ONE GEMM [voxels][Cin] x [Cin][taps*Cout] whose epilogue scatters column block t to output phase t
  Upscale up;
  up.sd = sd; up.sh = sh; up.sw = sw; up.cout = y->c;
  return conv_fprop_umma_v(view_of(x), w_packed, bias, view_of(y), 1, 1, 1, 0, (cudaStream_t)stream, up);
}

B200_EXPORT int b200_convT_dgrad_tc(const b200_tensor* dy, const void* w_packed_t, const b200_tensor* dx, int32_t sd,
                                    int32_t sh, int32_t sw, int32_t accumulate, void* stream) {
  B200_CHECK_ARG(check_tensor(dy, "convT_dgrad_tc.dy") && check_tensor(dx, "convT_dgrad_tc.dx") && w_packed_t, "%s",
                 b200_last_error());
  B200_CHECK_ARG(convT_tc_ok(dx, dy, sd, sh, sw), "convT_dgrad_tc: unsupported operands");
  const ActView xv = view_of(dx);
  ActView phases[8];
  const int T = sd * sh * sw;
  if (T <= 8) {
    int t = 0;
    for (int a = 0; a < sd; ++a)
      for (int b = 0; b < sh; ++b)
        for (int c = 0; c < sw; ++c, ++t) phases[t] = phase_view(dy, sd, sh, sw, a, b, c);
    // one GEMM with K = (phase, Cout): every K block reads its A box through the tensor map of its phase
    return conv_fprop_phases_v(phases, T, w_packed_t, xv, accumulate ? 1 : 0, (cudaStream_t)stream);
  }
  int t = 0;
  for (int a = 0; a < sd; ++a)
    for (int b = 0; b < sh; ++b)
      for (int c = 0; c < sw; ++c, ++t) {
        const ActView yv = phase_view(dy, sd, sh, sw, a, b, c);
        const char* wt = (const char*)w_packed_t + (size_t)t * dy->c * dx->c * 2;     // [Cin][Cout] for this phase
        int rc = conv_fprop_umma_v(yv, wt, nullptr, xv, 1, 1, 1, (accumulate || t > 0) ? 1 : 0, (cudaStream_t)stream);
        if (rc) return rc;
      }
  return B200_OK;
}

B200_EXPORT int b200_convT_wgrad_tc(const b200_tensor* x, const b200_tensor* dy, float* dw_packed, float* dbias, int32_t sd,
                                    int32_t sh, int32_t sw, void* stream) {
  B200_CHECK_ARG(check_tensor(x, "convT_wgrad_tc.x") && check_tensor(dy, "convT_wgrad_tc.dy") && dw_packed, "%s",
                 b200_last_error());
  B200_CHECK_ARG(convT_tc_ok(x, dy, sd, sh, sw), "convT_wgrad_tc: unsupported operands");
  const ActView xv = view_of(x);
  const int T = sd * sh * sw;
  if (T <= 8) {
    ActView phases[8];
    int t = 0;
    for (int a = 0; a < sd; ++a)
      for (int b = 0; b < sh; ++b)
        for (int c = 0; c < sw; ++c, ++t) phases[t] = phase_view(dy, sd, sh, sw, a, b, c);
    // GEMM N dimension = (phase, Cout): the dY boxes of up to 256 / Cout phases sit side by side in one B tile
    int rc = conv_wgrad_phases_v(xv, phases, T, dw_packed, (cudaStream_t)stream);
    if (rc) return rc;
  } else {
    int t = 0;
    for (int a = 0; a < sd; ++a)
      for (int b = 0; b < sh; ++b)
        for (int c = 0; c < sw; ++c, ++t) {
          const ActView yv = phase_view(dy, sd, sh, sw, a, b, c);
          int rc = conv_wgrad_umma_v(xv, yv, dw_packed + (size_t)t * dy->c * x->c, 1, 1, 1, (cudaStream_t)stream);
          if (rc) return rc;
        }
  }
  if (dbias) return conv_bias_grad(dy, dbias, (cudaStream_t)stream);
  return B200_OK;
}

// ------------------------------------------------------------------------------- tensor-pipe microbenchmark (diagnostic)
// Issues `iters` tcgen05.mma (M = 128, N = n, K = 16, bf16) from one thread per CTA over shared memory tiles and reports
// cycles per instruction: the ceiling the conv kernels' MMA lane can reach for a given N / swizzle / commit cadence.
namespace b200 { namespace sm100 {
template <int CE>
__global__ void __launch_bounds__(128, 1)
umma_rate_kernel(int n, int groups, uint32_t idesc, long long* out) {
  // CE = MMAs per commit (0: one commit at the very end); descriptors are precomputed so the issue loop is MMA + commit only
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_bar[9];
  __shared__ uint32_t s_tmem;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 9; ++i) mbar_init(smem_u32(&s_bar[i]), 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&s_tmem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (threadIdx.x == 0) {
    constexpr int G = CE > 0 ? CE : 8;
    uint64_t ad[4], bd[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ad[k] = make_smem_desc(smem0 + k * 32, 16, 1024, kSwizzle128);
      bd[k] = make_smem_desc(smem0 + 48 * 1024 + k * 32, 16, 1024, kSwizzle128);
    }
    const long long t0 = clock64();
    uint32_t ring = smem_u32(&s_bar[0]);
    for (int g = 0; g < groups; ++g) {
#pragma unroll
      for (int j = 0; j < G; ++j) umma_f16(tmem + ((j & 1) ? (uint32_t)n : 0u), ad[j & 3], bd[j & 3], idesc, 1u);
      if (CE > 0) {
        umma_commit(ring);
        ring = (g & 7) == 7 ? smem_u32(&s_bar[0]) : ring + 8;
      }
    }
    umma_commit(smem_u32(&s_bar[8]));
    const long long t_issue = clock64();
    mbar_wait(smem_u32(&s_bar[8]), 0);
    const long long t1 = clock64();
    out[2 * blockIdx.x] = t1 - t0;
    out[2 * blockIdx.x + 1] = t_issue - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// Handshake latency: thread 0 of warp 0 signals bar1 (plain arrive, or tcgen05.commit on an idle pipe), `nwait` threads of
// warps 1.. wait for it and arrive on bar2, thread 0 waits for bar2.  Reports cycles per round trip.
__global__ void __launch_bounds__(160, 1)
handshake_kernel(int iters, int use_commit, int use_test_wait, int nwait, long long* out) {
  __shared__ uint64_t s_bar[2];
  __shared__ uint32_t s_tmem;
  const uint32_t b1 = smem_u32(&s_bar[0]), b2 = smem_u32(&s_bar[1]);
  if (threadIdx.x == 0) {
    mbar_init(b1, 1);
    mbar_init(b2, nwait);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&s_tmem), 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  auto wait = [&](uint32_t bar, uint32_t par) {
    if (use_test_wait) { while (!mbar_test_wait(bar, par)) {} }
    else { while (!mbar_try_wait(bar, par)) {} }
  };
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (use_commit) umma_commit(b1); else mbar_arrive(b1);
      wait(b2, i & 1);
    }
    out[blockIdx.x] = clock64() - t0;
  } else if (threadIdx.x >= 32 && (int)threadIdx.x < 32 + nwait) {
    for (int i = 0; i < iters; ++i) {
      wait(b1, i & 1);
      mbar_arrive(b2);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(s_tmem, 32); }
}

// TMA rate: each CTA streams `iters` boxes of [box_rows][inner] 16-bit elements of an L2-resident matrix through a ring of
// `stages` shared-memory slots; reports cycles per box (one elected thread issues and waits, no consumer work).
__global__ void __launch_bounds__(128, 1)
tma_rate_kernel(const __grid_constant__ CUtensorMap tm, int iters, int stages, int box_rows, int box_bytes, int total_rows,
                long long* out) {
  // every warp's lane 0 drives an independent ring (blockDim.x / 32 concurrent issuers)
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_bar[4][8];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&s_bar[warp][i]), 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm);
    const uint32_t slot = (uint32_t)(box_rows * box_bytes + 1023) & ~1023u;
    const uint32_t base = smem0 + warp * stages * slot;
    const int nboxes = total_rows / box_rows;
    int row_box = ((blockIdx.x * 4 + warp) * 37) % nboxes;
    const long long t0 = clock64();
    for (int i = 0; i < iters + stages; ++i) {
      const int s = i % stages;
      if (i >= stages) mbar_wait(smem_u32(&s_bar[warp][s]), ((i / stages) - 1) & 1);
      if (i < iters) {
        mbar_expect_tx(smem_u32(&s_bar[warp][s]), (uint32_t)(box_rows * box_bytes));
        tma_load_2d(base + s * slot, &tm, smem_u32(&s_bar[warp][s]), 0, row_box * box_rows);
        row_box += 1;
        if (row_box >= nboxes) row_box = 0;
      }
    }
    out[blockIdx.x * 4 + warp] = clock64() - t0;
  }
}
}}  // namespace b200::sm100

static int tma_rate_sweep(int verbose, cudaStream_t st, long long* d_out) {
  using namespace b200;
  using namespace b200::sm100;
  const int total_rows = 256 * 1024;                 // x 128 B = 32 MB: L2 resident after the first pass
  void* buf = nullptr;
  B200_CUDA(cudaMalloc(&buf, (size_t)total_rows * 128));
  B200_CUDA(cudaMemsetAsync(buf, 0, (size_t)total_rows * 128, st));
  auto kern = tma_rate_kernel;
  B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int iters = 2048;
  for (int inner : {16, 64})
    for (int rows : {64, 128, 256})
      for (int warps : {1, 2, 4})
        for (int stages : {2, 4}) {
          if ((size_t)warps * stages * rows * inner * 2 > 190 * 1024) continue;
          CUtensorMap tm;
          int rc = make_matrix_tmap(&tm, buf, B200_BF16, total_rows, 64, rows, inner);
          if (rc) return rc;
          for (int rep = 0; rep < 2; ++rep)
            kern<<<148, 32 * warps, 200 * 1024, st>>>(tm, iters, stages, rows, inner * 2, total_rows, d_out);
          B200_LAUNCH_CHECK();
          long long h[148 * 4];
          B200_CUDA(cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, st));
          B200_CUDA(cudaStreamSynchronize(st));
          long long worst = 0;
          for (int b = 0; b < 148; ++b)
            for (int w = 0; w < warps; ++w) if (h[b * 4 + w] > worst) worst = h[b * 4 + w];
          const double cyc = (double)worst / (iters * warps);
          if (verbose)
            printf("tma_rate box=%dx%dB issuers=%d stages=%d: %.0f cycles/box/SM, %.1f B/clk/SM\n", rows, inner * 2, warps, stages,
                   cyc, rows * inner * 2 / cyc);
        }
  cudaFree(buf);
  return B200_OK;
}

B200_EXPORT int b200_umma_selftest(int32_t verbose, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  long long* d_out = nullptr;
  const int max_blocks = 148;
  B200_CUDA(cudaMalloc(&d_out, sizeof(long long) * 4 * max_blocks));
  for (int use_commit = 0; use_commit < 2; ++use_commit)
    for (int use_test = 0; use_test < 2; ++use_test)
      for (int nwait : {1, 32, 128}) {
        handshake_kernel<<<148, 160, 0, st>>>(4096, use_commit, use_test, nwait, d_out);
        B200_LAUNCH_CHECK();
        long long h[148];
        B200_CUDA(cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaStreamSynchronize(st));
        long long worst = 0;
        for (int b = 0; b < 148; ++b) if (h[b] > worst) worst = h[b];
        if (verbose)
          printf("handshake signal=%s wait=%s waiters=%d: %.0f cycles per round trip\n", use_commit ? "tcgen05.commit" : "arrive",
                 use_test ? "test_wait" : "try_wait", nwait, (double)worst / 4096);
      }
  if (verbose > 1) {
    int rc = tma_rate_sweep(verbose, st, d_out);
    if (rc) return rc;
  }
  if (verbose > 1) {
    typedef void (*RateKern)(int, int, uint32_t, long long*);
    const RateKern kerns[7] = {umma_rate_kernel<0>, umma_rate_kernel<1>, umma_rate_kernel<2>, umma_rate_kernel<4>,
                               umma_rate_kernel<8>, umma_rate_kernel<16>, umma_rate_kernel<32>};
    const int ces[7] = {0, 1, 2, 4, 8, 16, 32};
    const int ns[4] = {64, 128, 192, 256};
    for (int ki = 0; ki < 7; ++ki) {
      B200_CUDA(cudaFuncSetAttribute(kerns[ki], cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      for (int ni = 0; ni < 4; ++ni) {
        const int per_group = ces[ki] > 0 ? ces[ki] : 8;
        const int groups = 8192 / per_group;
        const uint32_t idesc = make_idesc(1, ns[ni], 0, 0);
        kerns[ki]<<<max_blocks, 128, 100 * 1024, st>>>(ns[ni], groups, idesc, d_out);
        B200_LAUNCH_CHECK();
        long long h[2 * max_blocks];
        B200_CUDA(cudaMemcpyAsync(h, d_out, sizeof(long long) * 2 * max_blocks, cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaStreamSynchronize(st));
        long long worst = 0, issue = 0;
        for (int b = 0; b < max_blocks; ++b) { if (h[2 * b] > worst) worst = h[2 * b]; if (h[2 * b + 1] > issue) issue = h[2 * b + 1]; }
        if (verbose)
          printf("umma_rate N=%d mma_per_commit=%d: %.1f cycles/MMA (issue %.1f), ideal %.0f\n", ns[ni], ces[ki],
                 (double)worst / 8192, (double)issue / 8192, ns[ni] / 2.0);
      }
    }
  }
  cudaFree(d_out);
  if (verbose) fflush(stdout);
  return B200_OK;
}
