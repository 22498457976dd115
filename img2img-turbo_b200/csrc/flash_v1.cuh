// flash_attn_v1_kernel: the round-1 softmax scheme (S read twice, PV tile read every step, O in registers), kept selectable with
// I2IT_FLASH_V1=1 for A/B measurements against flash_attn_kernel (single S read, O in TMEM with lazy rescale).  Same parameters,
// same tensor maps, same launch shape.
#pragma once
#include "flash.cuh"

namespace i2it {

template <typename T>
__global__ void __launch_bounds__(FA_THREADS, 2)
flash_attn_v1_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmVt, const __grid_constant__ FlashParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base;
  const uint32_t sKV = base + FA_Q_BYTES;
  const uint32_t sP = sKV + FA_STAGES * FA_KV_STAGE;
  const uint32_t bars = sP + FA_P_BYTES;
  const uint32_t q_full = bars;
  auto kv_full = [&](int s) { return bars + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bars + 8u * (1 + FA_STAGES + s); };
  // S is double-buffered in TMEM; each buffer has its own full/free barrier so no waiter can fall two phases behind
  const uint32_t s_full0 = bars + 8u * (1 + 2 * FA_STAGES);
  auto s_full = [&](int u) { return s_full0 + 8u * u; };
  auto s_free = [&](int u) { return s_full0 + 16u + 8u * u; };
  const uint32_t p_full = s_full0 + 32, pv_full = s_full0 + 40;
  const uint32_t tmem_slot = s_full0 + 48;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x % p.q_tiles;
  const int h = (blockIdx.x / p.q_tiles) % p.heads;
  const int b = blockIdx.x / (p.q_tiles * p.heads);
  int nkv = (p.Nk + FA_BN - 1) / FA_BN;
  if (p.causal) nkv = min(nkv, (qt * FA_BM + FA_BM + FA_BN - 1) / FA_BN);   // same value in all three roles

  if (warp == 4 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < FA_STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int u = 0; u < 2; ++u) { mbar_init(s_full(u), 1); mbar_init(s_free(u), 4); }
    mbar_init(p_full, 4); mbar_init(pv_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmK)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmVt)) : "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_sync();   // prologue (barriers, TMEM, descriptor prefetch) overlaps the previous kernel's tail; no global access before here
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  const uint32_t tS0 = tmem_base, tPV = tmem_base + 128;   // S buffers at columns [0,64) and [64,128)

  if (warp == 4) {
    // producer: warp-uniform loop, one elected lane issues the TMA
    if (elect_one()) {
      mbar_expect_tx(q_full, FA_Q_BYTES);
      tma_load_5d(sQ, &tmQ, q_full, 0, qt * FA_BM, h, b, 0);
    }
    __syncwarp();
    for (int j = 0; j < nkv; ++j) {
      const int s = j % FA_STAGES;
      mbar_wait(kv_empty(s), ((j / FA_STAGES) & 1) ^ 1, p.err, 11);
      if (elect_one()) {
        mbar_expect_tx(kv_full(s), FA_KV_DATA);      // (the stage stride also holds the current kernel's 2 KB of constant rows)
        tma_load_5d(sKV + s * FA_KV_STAGE, &tmK, kv_full(s), 0, j * FA_BN, h, b * p.kv_bmul, 0);
        tma_load_5d(sKV + s * FA_KV_STAGE + FA_BN * FA_D * 2, &tmVt, kv_full(s), j * FA_BN, 0, h, b * p.kv_bmul, 0);
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    // MMA issuer: warp-uniform loop, one elected lane issues tcgen05.mma / commit
    mbar_wait(q_full, 0, p.err, 12);
    const uint64_t qdesc = umma_desc_sw128(sQ), pdesc = umma_desc_sw128(sP);
    auto issue_qk = [&](int j) {
      const int s = j % FA_STAGES;
      mbar_wait(kv_full(s), (j / FA_STAGES) & 1, p.err, 13);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t kdesc = umma_desc_sw128(sKV + s * FA_KV_STAGE);
#pragma unroll
        for (int k = 0; k < FA_D / 16; ++k)
          tc_mma_f16(tS0 + 64 * (j & 1), qdesc + 2 * k, kdesc + 2 * k, p.idesc, k > 0 ? 1u : 0u);
        tc_commit(s_full(j & 1));
      }
      __syncwarp();
    };
    issue_qk(0);
    for (int j = 0; j < nkv; ++j) {
      if (j + 1 < nkv) {
        // S buffer (j+1)&1 was last read by softmax iteration j-1: QK_{j+1} overlaps the exponentials of iteration j
        if (j >= 1) { mbar_wait(s_free((j + 1) & 1), ((j - 1) >> 1) & 1, p.err, 14); tc_fence_after(); }
        issue_qk(j + 1);
      }
      mbar_wait(p_full, j & 1, p.err, 15);        // P_j is in smem (and PV_{j-1} has been consumed)
      tc_fence_after();
      const int s = j % FA_STAGES;
      if (elect_one()) {
        const uint64_t vdesc = umma_desc_sw128(sKV + s * FA_KV_STAGE + FA_BN * FA_D * 2);
#pragma unroll
        for (int k = 0; k < FA_BN / 16; ++k) tc_mma_f16(tPV, pdesc + 2 * k, vdesc + 2 * k, p.idesc, k > 0 ? 1u : 0u);
        tc_commit(pv_full);
        tc_commit(kv_empty(s));                     // K_j and V_j are free once everything issued so far retires
      }
      __syncwarp();
    }
  } else {
    // ---------------- softmax / output warps: thread = Q row ----------------
    const int row = warp * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    float o[FA_D];
#pragma unroll
    for (int i = 0; i < FA_D; ++i) o[i] = 0.f;
    float m = -INFINITY, l = 0.f, alpha_prev = 1.f;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full(j & 1), (j >> 1) & 1, p.err, 16);
      tc_fence_after();
      const uint32_t tS = tS0 + 64 * (j & 1);
      const int kbase = j * FA_BN;
      const int qpos = qt * FA_BM + row;
      // element-wise masking only where needed (warp-uniform): the last KV tile, or causal tiles that reach this warp's diagonal
      const bool ragged = (kbase + FA_BN > p.Nk) || (p.causal && kbase + FA_BN - 1 > qt * FA_BM + warp * 32);
      const int klim = p.causal ? min(p.Nk, qpos + 1) : p.Nk;   // keys [0, klim) are visible to this row
      const float sc = p.scale_log2e;
      // pass 1: row max of the RAW logits (scale > 0 keeps the order; one FMNMX per element)
      float mx = -INFINITY;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t raw[32];
        tc_ld32(tS + lane_off + half * 32, raw);
        tc_wait_ld();
        if (ragged) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (kbase + half * 32 + i < klim) mx = fmaxf(mx, __uint_as_float(raw[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(raw[i]));
        }
      }
      const float m_new = fmaxf(m, mx);                    // m, m_new in raw (unscaled) units
      const float alpha = fast_exp2((m - m_new) * sc);     // first tile: exp2(-inf) = 0
      const float neg_ms = -m_new * sc;
      // fold in PV_{j-1} before P_{j-1}'s smem tile is overwritten
      if (j > 0) {
        mbar_wait(pv_full, (j - 1) & 1, p.err, 17);
        tc_fence_after();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t raw[32];
          tc_ld32(tPV + lane_off + half * 32, raw);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[half * 32 + i] = o[half * 32 + i] * alpha_prev + __uint_as_float(raw[i]);
        }
      }
      // pass 2: p = exp2(s*scale - m*scale) (one FFMA + one MUFU per element), fp32 row sum of the unrounded p (as
      // FlashAttention does), pack to 16 bit, write the swizzled K-major P tile
      float psum = 0.f;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t raw[32];
        tc_ld32(tS + lane_off + half * 32, raw);
        tc_wait_ld();
        uint32_t pk[16];
        if (ragged) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int k0 = kbase + half * 32 + 2 * i;
            const float p0 = (k0 < klim) ? fast_exp2(fmaf(__uint_as_float(raw[2 * i]), sc, neg_ms)) : 0.f;
            const float p1 = (k0 + 1 < klim) ? fast_exp2(fmaf(__uint_as_float(raw[2 * i + 1]), sc, neg_ms)) : 0.f;
            psum += p0 + p1;
            pk[i] = Elem<T>::pack(p0, p1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = fast_exp2(fmaf(__uint_as_float(raw[2 * i]), sc, neg_ms));
            const float p1 = fast_exp2(fmaf(__uint_as_float(raw[2 * i + 1]), sc, neg_ms));
            psum += p0 + p1;
            pk[i] = Elem<T>::pack(p0, p1);
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {               // 16-byte group g' = half*4+g holds keys 8g'..8g'+7 of this row
          const int gg = half * 4 + g;
          sts16(sP + row * 128 + ((gg ^ (row & 7)) << 4), pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
        }
      }
      l = l * alpha + psum;
      m = m_new;
      alpha_prev = alpha;
      // S_j fully consumed -> the MMA warp may overwrite S with QK_{j+1};  P_j visible to the async proxy -> PV_j may start
      tc_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) { mbar_arrive(s_free(j & 1)); mbar_arrive(p_full); }
    }
    // last PV
    mbar_wait(pv_full, (nkv - 1) & 1, p.err, 18);
    tc_fence_after();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t raw[32];
      tc_ld32(tPV + lane_off + half * 32, raw);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[half * 32 + i] = o[half * 32 + i] * alpha_prev + __uint_as_float(raw[i]);
    }
    const int q = qt * FA_BM + row;
    if (q < p.Nq) {
      const float inv = 1.0f / l;
      T* optr = reinterpret_cast<T*>(p.out) + (static_cast<long long>(b) * p.Nq + q) * p.ldo + h * FA_D;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint4 u;
        u.x = Elem<T>::pack(o[8 * g] * inv, o[8 * g + 1] * inv);
        u.y = Elem<T>::pack(o[8 * g + 2] * inv, o[8 * g + 3] * inv);
        u.z = Elem<T>::pack(o[8 * g + 4] * inv, o[8 * g + 5] * inv);
        u.w = Elem<T>::pack(o[8 * g + 6] * inv, o[8 * g + 7] * inv);
        st16(optr + 8 * g, u);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
  }
}

}  // namespace i2it
