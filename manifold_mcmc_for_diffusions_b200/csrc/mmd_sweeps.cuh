// Time-stepping sweeps and per-block solves shared by the CHMC phases (mmd_kernels_main.cuh).
// Everything here is per thread = per (chain, observation block); see mmd_kernels.cuh for the layout.
#pragma once
#include "mmd_kernels.cuh"

// the per-iteration block solves: fully unrolled (factor loads issued up front: lowest latency for a lone
// CTA) or rolled (fewer registers: higher throughput with all SMs busy)
#if defined(MMD_ROLLED_SOLVES)
#define MMD_SOLVE_UNROLL _Pragma("unroll 1")
#else
#define MMD_SOLVE_UNROLL _Pragma("unroll")
#endif

namespace mmd {

// per-thread parameters of the chain at some value of u
template <class M, int UMAX>
struct ChainPar {
  double u[UMAX];
  double z[M::Z];
  double dzdu[M::Z * M::Z];
  typename M::Coef C;
  double sigy;
};
template <class M, int UMAX>
MMD_D void make_par(const Dims& d, const double* u, ChainPar<M, UMAX>& P) {
#pragma unroll
  for (int j = 0; j < UMAX; ++j) P.u[j] = u[j];
  M::gen_z(d.gen, P.u, P.z, P.dzdu);
  M::make_coef(P.z, d.sd, P.C);
  P.sigy = sigma_of<M>(d, P.u);
}

// state at the start of the thread's block: generate_x_0(z, v_0) for the first block, otherwise the
// conditioned state at the end of the previous block (partition_into_subseqs, :413-471)
template <class M>
MMD_D void block_start(const Dims& d, const Blk& B, const double* z, const double* v0, const double* xoc, int cpb,
                       double* x) {
  if (B.ini) {
    M::gen_x0(d.gen, z, v0, x);
  } else {
    ldcol<M::X>(xoc + (B.o - 1) * M::X * cpb, cpb, x);
  }
}

// ------------------------------------------------------------------------------------------
// asynchronous global -> shared copies (LDGSTS): the sweeps prefetch the records of future time steps
// into a thread-private shared-memory ring, so the recursion never waits on HBM and no registers are
// spent on data in flight.  Shared addresses are 32-bit window offsets computed once per sweep.
// ------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
MMD_D unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// L2 residency hints.  The per-iteration block factors (L, A, D^-1 A, Psi_k, kappa_k, alpha_k: ~100 doubles per
// thread) are re-read by every solver iteration while the streams (work position, compressed Jacobian: 6 KB per
// thread and sweep) flush the L2 in between; evict_last keeps the factors resident, evict_first marks the streams
// as the first candidates for replacement.
MMD_D unsigned long long l2_policy_keep() {
  unsigned long long p;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
MMD_D unsigned long long l2_policy_stream() {
  unsigned long long p;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
MMD_D double ldg_keep(const double* g) {
#if defined(MMD_HINT_KEEP)
  double v;
  asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;\n" : "=d"(v) : "l"(g), "l"(l2_policy_keep()));
  return v;
#else
  return *g;
#endif
}
// The 16-byte copies allocate in L1 (.ca, not the L1-bypassing .cg): a thread's X*V-double record is fetched in
// 16-byte pieces at a lane stride of X*V*8 bytes, so every copy instruction of a warp touches half of each 32-byte
// sector and the next one the other half.  With .cg both requests go to L2 and the sweeps saturate the L2 request
// path at 4.5 TB/s of useful bytes; with .ca the second one hits L1: 6.9 TB/s, next to the 7.2 TB/s of CTA-wide
// cp.async.bulk copies (tools/stream_bench.cu, profiles/r2_stream_bench.json).
#if defined(MMD_CP_ASYNC_CG)
#define MMD_CP16 "cp.async.cg"
#else
#define MMD_CP16 "cp.async.ca"
#endif
template <int BYTES>
MMD_D void cp_async(unsigned sdst, const void* gsrc) {
#if defined(MMD_HINT_STREAM)
  if (BYTES == 16)
    asm volatile(MMD_CP16 ".shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(sdst), "l"(gsrc), "l"(l2_policy_stream()) : "memory");
  else
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;\n" ::"r"(sdst), "l"(gsrc), "l"(l2_policy_stream()) : "memory");
#else
  if (BYTES == 16)
    asm volatile(MMD_CP16 ".shared.global [%0], [%1], 16;\n" ::"r"(sdst), "l"(gsrc) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sdst), "l"(gsrc) : "memory");
#endif
}
// plain volatile load: consecutive calls stay together and in order, so a batch of independent loads is in flight at
// once instead of each one being scheduled next to its use
MMD_D double ldg_vol(const double* g) {
  double v;
#if defined(MMD_KEEP_FACTORS)
  // evict_last: the factors are re-read by every solver iteration while ~350 MB of streams pass through L2 in between
  asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;\n" : "=d"(v) : "l"(g), "l"(l2_policy_keep()));
#else
  asm volatile("ld.global.f64 %0, [%1];\n" : "=d"(v) : "l"(g));
#endif
  return v;
}
template <int N>
MMD_D void ldcol_keep(const double* g, int ld, double* r) {  // ldcol with the evict_last hint
#pragma unroll
  for (int i = 0; i < N; ++i) r[i] = ldg_keep(g + i * ld);
}
MMD_D void prefetch_l2(const void* g) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(g)); }
MMD_D void prefetch_l1(const void* g) { asm volatile("prefetch.global.L1 [%0];\n" ::"l"(g)); }
// prefetch `rows` rows of a thread-private column into L2 (the demand loads of the block solve that follows the
// sweep then hit L2 instead of queueing behind the sweeps' streams in DRAM)
MMD_D void prefetch_col_l2(const double* g, int rows, int ld) {
  for (int i = 0; i < rows; ++i) prefetch_l2(g + i * ld);
}
MMD_D void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
MMD_D void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
// ------------------------------------------------------------------------------------------
// Blackwell bulk copies: one elected lane per warp moves the warp's contiguous slab of a [step][nta][W] array
// with cp.async.bulk (the TMA engine: no per-thread address arithmetic, no LDGSTS issue, no registers for data
// in flight); completion is counted in bytes on an mbarrier the lanes of the warp wait on.
// ------------------------------------------------------------------------------------------
MMD_D void mbar_init(unsigned bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
MMD_D void mbar_inval(unsigned bar) { asm volatile("mbarrier.inval.shared::cta.b64 [%0];\n" ::"r"(bar) : "memory"); }
MMD_D void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
MMD_D void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MMD_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MMD_DONE_%=;\n"
      "bra MMD_WAIT_%=;\n"
      "MMD_DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
MMD_D void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
#if defined(MMD_HINT_STREAM)
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar), "l"(l2_policy_stream()) : "memory");
#else
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar) : "memory");
#endif
}
MMD_D void bulk_prefetch_l2(const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(src), "r"(bytes) : "memory");
}
MMD_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
// copy an N-double record (16-byte pieces when N is even)
template <int N>
MMD_D void cp_async_rec(unsigned sdst, const double* gsrc) {
  if (N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i) cp_async<16>(sdst + 16 * i, gsrc + 2 * i);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) cp_async<8>(sdst + 8 * i, gsrc + i);
  }
}
template <int N>
MMD_D void lds_rec(unsigned saddr, double* r) {
  if (N % 2 == 0) {
#pragma unroll
    for (int i = 0; i < N / 2; ++i)
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(r[2 * i]), "=d"(r[2 * i + 1]) : "r"(saddr + 16 * i) : "memory");
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(r[i]) : "r"(saddr + 8 * i) : "memory");
  }
}
#endif

// ring record per thread and slot: [v (V) | K (X*V)], padded to an even number of doubles
template <class M>
struct RingRec {
  static constexpr int NW = M::V + M::X * M::V;
  static constexpr int NWP = (NW + 1) / 2 * 2;
};

// ------------------------------------------------------------------------------------------
// forward constraint sweep for one block:  c_b(q)   (generate_y_bar + constr, :399-411, :473-519).
// With WITH_K the position is the quasi-Newton parametrisation q = qw - J_lin^T lambda_tot, never
// materialised:   v_t = qw_v[t] - K_t^T alpha_k ,  n_k = qw_n[k] - sigma_lin * lambda_tot[k].
// The records of step s + PF are fetched by cp.async into a shared-memory ring while step s is
// computed (the loads do not depend on the recursion; the sweep would otherwise be bound by
// global-load latency).  ring: CTA ring base; slot i of thread tid at ring[(i * NT + tid) * NWP].
// crow / lamtot are thread-private columns in shared memory: element r at [r * NT].
// Not inlined on purpose: the loop gets its own register allocation, independent of what the caller
// keeps alive across the sweep.
// ------------------------------------------------------------------------------------------
template <class M>
struct SweepArgs {
  typename M::Coef C;
  double sigma_y, sigma_lin;
  double xstart[M::X];
  const double* vb;      // q-like body records of the block (V per step)
  const double* nzb;     // noise column
  const double* xoc;     // per-chain x_obs_seq column
  const double* y;
  const double* Kb;      // compressed Jacobian records (X*V per step)
  const double* alph;    // thread-private alpha column [rmax*X]
  const double* lamtot;  // shared-memory column
  double* crow;          // shared-memory column (out)
  double* ring;          // CTA ring base (shared)
  double* xend_out;      // thread-private [rmax*X] or null
  double* xs_out;        // thread-private trajectory records [rmax*S][X] (state BEFORE each step) or null
  int nta, cpb, NT, tid;
  unsigned mask;         // lanes of this warp that run the sweep (0: caller did not form it -> per-thread cp.async path)
  unsigned long long* bars;  // one mbarrier per warp (shared memory)
  unsigned long long* phase;  // MMD_PHASE_CLOCK builds: cycle counters (slot 31: waits for the ring, slot 8: recursion)
};

// Warp-cooperative form of the quasi-Newton sweep on bulk copies.  The ring holds NSL steps: slot u =
// [v slab: NT x V][K slab: NT x X*V] exactly as the arrays lie in global memory, so the records of one warp are one
// contiguous piece per array and step.  Per group of NSL steps: wait on the warp's mbarrier, read the NSL records
// and form v_t = qw_t - K_t^T alpha_k (independent of the recursion), __syncwarp, the elected lane re-arms the
// barrier and issues the 2 * NSL bulk copies of the next group (plus an L2 prefetch of the one after), then the
// NSL serial steps run while the copies are in flight.  Lanes whose block has fewer intervals stay in the loop
// (predicated) so the warp keeps its barrier protocol; lanes outside `mask` never enter.  Same floating-point
// operations in the same order as constr_sweep.
template <class M, bool WITH_K>
__device__ __noinline__ void constr_sweep_bulk(const Dims& d, const Blk& B, const SweepArgs<M>& a) {
  constexpr int X = M::X, V = M::V, XV = M::X * M::V;
  constexpr int NSL = MMD_PREFETCH_STEPS + 1;
  const int nta = a.nta, NT = a.NT, S = d.S;
  const unsigned mask = a.mask;
  const int lane = a.tid & 31, warp = a.tid >> 5;
  const bool leader = lane == (__ffs(mask) - 1);
  const int nl = (NT - warp * 32) < 32 ? (NT - warp * 32) : 32;     // threads of this warp
  const int nkw = __reduce_max_sync(mask, B.n);                      // intervals the warp walks through
  const int nsw = nkw * S;
  const typename M::Coef C = a.C;
  double x[X];
#pragma unroll
  for (int i = 0; i < X; ++i) x[i] = a.xstart[i];
  // shared-memory addresses: slot u at ring + u * slot_bytes; inside a slot the v slab, then the K slab
  const unsigned slot_bytes = (unsigned)NT * (RingRec<M>::NWP * 8);
  const unsigned ring = smem_u32(a.ring);
  const unsigned my_v = ring + (unsigned)a.tid * (V * 8);
  const unsigned my_K = ring + (unsigned)NT * (V * 8) + (unsigned)a.tid * (XV * 8);
  const unsigned w_v = ring + (unsigned)(warp * 32) * (V * 8);
  const unsigned w_K = ring + (unsigned)NT * (V * 8) + (unsigned)(warp * 32) * (XV * 8);
  const unsigned bar = smem_u32(a.bars) + 8u * warp;
  const double* vw = a.vb - lane * V;     // the warp's slab of step 0
  const double* Kw = WITH_K ? a.Kb - lane * XV : nullptr;
  const unsigned vbytes = (unsigned)nl * (V * 8), Kbytes = WITH_K ? (unsigned)nl * (XV * 8) : 0u;
  const int vbump = V * nta, Kbump = XV * nta;
  auto issue_group = [&](int s0) {   // leader only: steps s0 .. s0 + NSL - 1 into slots 0 .. NSL - 1
    mbar_expect_tx(bar, NSL * (vbytes + Kbytes));
#pragma unroll
    for (int u = 0; u < NSL; ++u) {
      bulk_g2s(w_v + u * slot_bytes, vw + (size_t)(s0 + u) * vbump, vbytes, bar);
      if (WITH_K) bulk_g2s(w_K + u * slot_bytes, Kw + (size_t)(s0 + u) * Kbump, Kbytes, bar);
    }
#if MMD_L2_PREFETCH_STEPS > 0
    if (s0 + 2 * NSL <= nsw - NSL) {   // the group after the next one: HBM -> L2
#pragma unroll
      for (int u = 0; u < NSL; ++u) {
        bulk_prefetch_l2(vw + (size_t)(s0 + 2 * NSL + u) * vbump, vbytes);
        if (WITH_K) bulk_prefetch_l2(Kw + (size_t)(s0 + 2 * NSL + u) * Kbump, Kbytes);
      }
    }
#endif
  };
#ifdef MMD_BULK_DEBUG
  if (blockIdx.x == 0) printf("enter tid %d warp %d lane %d mask %x leader %d nl %d nkw %d S %d vb %u Kb %u bar %u ring %u NT %d nta %d Bn %d\n", a.tid, warp, lane, mask, (int)leader, nl, nkw, S, vbytes, Kbytes, bar, ring, NT, nta, B.n);
#endif
  if (leader) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    fence_proxy_async();   // the ring region was last touched through the generic proxy (reduction scratch)
#if MMD_L2_PREFETCH_STEPS > 0
    if (2 * NSL <= nsw) {
#pragma unroll
      for (int u = 0; u < NSL; ++u) {
        bulk_prefetch_l2(vw + (size_t)(NSL + u) * vbump, vbytes);
        if (WITH_K) bulk_prefetch_l2(Kw + (size_t)(NSL + u) * Kbump, Kbytes);
      }
    }
#endif
    issue_group(0);
  }
  __syncwarp(mask);
  unsigned parity = 0;
  const int gpi = S / NSL;
  double al[X], aln[X];
#pragma unroll
  for (int i = 0; i < X; ++i) al[i] = aln[i] = 0.0;
  if (WITH_K) ldcol<X>(a.alph, nta, al);
  int s0 = 0;
  for (int k = 0; k < nkw; ++k) {
    const bool live = k < B.n;
    double yk = 0.0, nk = 0.0;
    if (live && k < B.ny) {
      yk = a.y[B.o + k];
      if (d.noisy) nk = a.nzb[k * nta];
    }
    if (WITH_K && k + 1 < B.n) ldcol<X>(a.alph + (k + 1) * X * nta, nta, aln);
    for (int g = 0; g < gpi; ++g) {
#ifdef MMD_PHASE_CLOCK
      const long long tw0 = clock64();
#endif
#ifdef MMD_BULK_DEBUG
      if (blockIdx.x == 0) printf("wait tid %d k %d g %d parity %u\n", a.tid, k, g, parity);
#endif
      mbar_wait(bar, parity);
      parity ^= 1u;
#ifdef MMD_BULK_DEBUG
      if (blockIdx.x == 0) printf("pass tid %d k %d g %d\n", a.tid, k, g);
#endif
#ifdef MMD_PHASE_CLOCK
      const long long tw1 = clock64();
      if (threadIdx.x == 0 && a.phase) atomicAdd(&a.phase[31], (unsigned long long)(tw1 - tw0));
#endif
      double v[NSL][V];
#pragma unroll
      for (int u = 0; u < NSL; ++u) {
        lds_rec<V>(my_v + u * slot_bytes, v[u]);
        if (WITH_K) {
          double Kt[XV];
          lds_rec<XV>(my_K + u * slot_bytes, Kt);
#pragma unroll
          for (int j = 0; j < V; ++j)
#pragma unroll
            for (int i = 0; i < X; ++i) v[u][j] = fma(-Kt[i * V + j], al[i], v[u][j]);
        }
      }
      s0 += NSL;
      __syncwarp(mask);   // every lane has read its records: the slots may be overwritten
      if (leader && s0 < nsw) issue_group(s0);
#ifdef MMD_PHASE_CLOCK
      const long long tc0 = clock64();
#endif
      if (live) {
#pragma unroll
        for (int u = 0; u < NSL; ++u) {
          double xn[X];
          M::step(C, x, v[u], xn);
#pragma unroll
          for (int i = 0; i < X; ++i) x[i] = xn[i];
        }
      }
#ifdef MMD_PHASE_CLOCK
      if (threadIdx.x == 0 && a.phase) {
        const long long tc1 = clock64();
        atomicAdd(&a.phase[8], (unsigned long long)(tc1 - tc0) + (unsigned long long)(x[0] == 1.2345e300 ? 1 : 0));
      }
#endif
    }
    if (live) {
      if (k < B.ny) {
        double cy = M::obs(x) - yk;
        if (d.noisy) {
          if (WITH_K) nk = fma(-a.sigma_lin, a.lamtot[k * NT], nk);
          cy = fma(a.sigma_y, nk, cy);
        }
        a.crow[k * NT] = cy;
      }
      if (k == B.n - 1 && B.nx > 0) {
        double xo[X];
        ldcol<X>(a.xoc + (B.o + k) * X * a.cpb, a.cpb, xo);
#pragma unroll
        for (int i = 0; i < X; ++i) a.crow[(B.ny + i) * NT] = x[i] - xo[i];
      }
    }
    if (WITH_K) {
#pragma unroll
      for (int i = 0; i < X; ++i) al[i] = aln[i];
    }
  }
  // the barrier object must be invalidated before the next sweep initialises it again
  __syncwarp(mask);
  if (leader) mbar_inval(bar);
}

template <class M, bool WITH_K, bool WITH_XS = false>
__device__ __noinline__ void constr_sweep(const Dims& d, const Blk& B, const SweepArgs<M>& a) {
  constexpr int X = M::X, V = M::V, XV = M::X * M::V;
  constexpr int PF = MMD_PREFETCH_STEPS, NSL = PF + 1, NWP = RingRec<M>::NWP;
  const int ns = B.n * d.S, nta = a.nta, NT = a.NT, S = d.S;
  const typename M::Coef C = a.C;
  double x[X];
#pragma unroll
  for (int i = 0; i < X; ++i) x[i] = a.xstart[i];
  const unsigned ring0 = smem_u32(a.ring) + (unsigned)a.tid * (NWP * 8);
  const unsigned sstride = (unsigned)NT * (NWP * 8);
  const unsigned ring_end = ring0 + NSL * sstride;
  const double* vp = a.vb;   // next record to fetch
  const double* Kp = a.Kb;
  const int vbump = V * nta, Kbump = XV * nta;
#if defined(MMD_GROUPED_SWEEP) || defined(MMD_BULK_SWEEP)   // opt-in (build with -DMMD_PREFETCH_STEPS=4): measured slower under full load, see DESIGN.md
  // Grouped form (quasi-Newton sweeps, S a multiple of the ring size): the recursion x_{t+1} = f(x_t, v_t) is a chain
  // of ~9 dependent FP64 operations per step, and with one or two warps per scheduler nothing else hides their
  // latency -- the step-at-a-time loop below runs at ~560 cycles per step on an otherwise idle SM although the chain
  // itself is ~80.  Here NSL steps are handled together: their records are read from the ring and turned into
  // v_t = qw_t - K_t^T alpha_k (independent of x: full instruction-level parallelism), the ring is refilled for the
  // next group right away, and only then the NSL serial steps run, with nothing but arithmetic between them.  The
  // observation row, noise variable and next alpha of an interval are fetched at its start.  Same operations in the
  // same order as the loop below: bit-identical results.
#if defined(MMD_BULK_SWEEP)
  if (!WITH_XS && (S % NSL) == 0 && (V % 2) == 0 && (XV % 2) == 0 && a.mask != 0u) {
    // bulk copies need 16-byte aligned slabs (warp-uniform test: the slab of the warp's first lane and the step strides)
    unsigned long long al16 = reinterpret_cast<unsigned long long>(a.vb - (a.tid & 31) * V) | (unsigned long long)(V * nta * 8);
    if (WITH_K) al16 |= reinterpret_cast<unsigned long long>(a.Kb - (a.tid & 31) * XV) | (unsigned long long)(XV * nta * 8);
    if ((al16 & 15ull) == 0ull) {
      constr_sweep_bulk<M, WITH_K>(d, B, a);
      return;
    }
  }
#endif
  if (!WITH_XS && (S % NSL) == 0) {
    auto fetch_group = [&](int s0) {   // records of steps s0 .. s0 + NSL - 1 into slots 0 .. NSL - 1
#pragma unroll
      for (int u = 0; u < NSL; ++u) {
        cp_async_rec<V>(ring0 + u * sstride, vp + u * vbump);
        if (WITH_K) cp_async_rec<XV>(ring0 + u * sstride + V * 8, Kp + u * Kbump);
      }
#if MMD_L2_PREFETCH_STEPS > 0
      if (s0 + 3 * NSL <= ns) {   // pull the group after the next one from HBM into L2
#pragma unroll
        for (int u = 0; u < NSL; ++u) {
          prefetch_l2(vp + (2 * NSL + u) * vbump);
          if (WITH_K) prefetch_l2(Kp + (2 * NSL + u) * Kbump);
        }
      }
#endif
      vp += NSL * vbump;
      Kp += NSL * Kbump;
      cp_async_commit();
    };
#if MMD_L2_PREFETCH_STEPS > 0
    if (2 * NSL <= ns) {
#pragma unroll
      for (int u = 0; u < NSL; ++u) {
        prefetch_l2(vp + (NSL + u) * vbump);
        if (WITH_K) prefetch_l2(Kp + (NSL + u) * Kbump);
      }
    }
#endif
    fetch_group(0);
    const int gpi = S / NSL;
    double al[X], aln[X];
    if (WITH_K) ldcol<X>(a.alph, nta, al);
    int s0 = 0;
    for (int k = 0; k < B.n; ++k) {
      // per-interval operands, in flight while the interval is integrated
      double yk = 0.0, nk = 0.0;
      if (k < B.ny) {
        yk = a.y[B.o + k];
        if (d.noisy) nk = a.nzb[k * nta];
      }
      if (WITH_K && k + 1 < B.n) ldcol<X>(a.alph + (k + 1) * X * nta, nta, aln);
      for (int g = 0; g < gpi; ++g) {
#ifdef MMD_PHASE_CLOCK
        const long long tw0 = clock64();
#endif
        cp_async_wait<0>();
#ifdef MMD_PHASE_CLOCK
        const long long tw1 = clock64();
        if (threadIdx.x == 0 && a.phase) atomicAdd(&a.phase[31], (unsigned long long)(tw1 - tw0));
#endif
        double v[NSL][V];
#pragma unroll
        for (int u = 0; u < NSL; ++u) {
          lds_rec<V>(ring0 + u * sstride, v[u]);
          if (WITH_K) {
            double Kt[XV];
            lds_rec<XV>(ring0 + u * sstride + V * 8, Kt);
#pragma unroll
            for (int j = 0; j < V; ++j)
#pragma unroll
              for (int i = 0; i < X; ++i) v[u][j] = fma(-Kt[i * V + j], al[i], v[u][j]);
          }
        }
        s0 += NSL;
        if (s0 < ns) fetch_group(s0);
#ifdef MMD_PHASE_CLOCK
        const long long tc0 = clock64();
#endif
#pragma unroll
        for (int u = 0; u < NSL; ++u) {
          double xn[X];
          M::step(C, x, v[u], xn);
#pragma unroll
          for (int i = 0; i < X; ++i) x[i] = xn[i];
        }
#ifdef MMD_PHASE_CLOCK
        if (threadIdx.x == 0 && a.phase) {
          const long long tc1 = clock64();
          atomicAdd(&a.phase[8], (unsigned long long)(tc1 - tc0) + (unsigned long long)(x[0] == 1.2345e300 ? 1 : 0));
        }
#endif
      }
      if (k < B.ny) {
        double cy = M::obs(x) - yk;
        if (d.noisy) {
          if (WITH_K) nk = fma(-a.sigma_lin, a.lamtot[k * NT], nk);
          cy = fma(a.sigma_y, nk, cy);
        }
        a.crow[k * NT] = cy;
      }
      if (k == B.n - 1 && B.nx > 0) {
        double xo[X];
        ldcol<X>(a.xoc + (B.o + k) * X * a.cpb, a.cpb, xo);
#pragma unroll
        for (int i = 0; i < X; ++i) a.crow[(B.ny + i) * NT] = x[i] - xo[i];
      }
      if (WITH_K) {
#pragma unroll
        for (int i = 0; i < X; ++i) al[i] = aln[i];
      }
    }
    cp_async_wait<0>();
    return;
  }
#endif
  unsigned wr = ring0;
#pragma unroll
  for (int i = 0; i < PF; ++i) {
    if (i < ns) {
      cp_async_rec<V>(wr, vp);
      if (WITH_K) cp_async_rec<XV>(wr + V * 8, Kp);
      vp += vbump;
      Kp += Kbump;
    }
    cp_async_commit();
    wr += sstride;
  }
  unsigned rd = ring0;
  double al[X], aln[X];
#pragma unroll
  for (int i = 0; i < X; ++i) aln[i] = 0.0;
  if (WITH_K) ldcol<X>(a.alph, nta, al);
  int s = 0;
  for (int k = 0; k < B.n; ++k) {
    // operands of the interval's end (observation, noise variable, conditioned state, next alpha): requested now, so
    // the loads are in flight while the interval is integrated instead of stalling the recursion at its last step
    double yk = 0.0, nk = 0.0, xo[X];
#pragma unroll
    for (int i = 0; i < X; ++i) xo[i] = 0.0;
    if (k < B.ny) {
      yk = a.y[B.o + k];
      if (d.noisy) nk = a.nzb[k * nta];
    }
    if (k == B.n - 1 && B.nx > 0) ldcol<X>(a.xoc + (B.o + k) * X * a.cpb, a.cpb, xo);
    if (WITH_K && k + 1 < B.n) ldcol<X>(a.alph + (k + 1) * X * nta, nta, aln);
    for (int t = 0; t < S; ++t, ++s) {
      cp_async_wait<PF - 1>();
      double v[V], xn[X];
      if (WITH_XS && a.xs_out) strec<X>(a.xs_out + s * X * nta, x);   // (Newton only: keeps the test out of the loop)
      lds_rec<V>(rd, v);
      if (WITH_K) {
        double Kt[XV];
        lds_rec<XV>(rd + V * 8, Kt);
#pragma unroll
        for (int j = 0; j < V; ++j)
#pragma unroll
          for (int i = 0; i < X; ++i) v[j] = fma(-Kt[i * V + j], al[i], v[j]);
      }
      if (s + PF < ns) {  // refill the slot consumed one step ago
        cp_async_rec<V>(wr, vp);
        if (WITH_K) cp_async_rec<XV>(wr + V * 8, Kp);
#if MMD_L2_PREFETCH_STEPS > 0
        if (s + PF + MMD_L2_PREFETCH_STEPS < ns) {  // pull the records of a later step from HBM into L2
          prefetch_l2(vp + MMD_L2_PREFETCH_STEPS * vbump);
          if (WITH_K) prefetch_l2(Kp + MMD_L2_PREFETCH_STEPS * Kbump);
        }
#endif
        vp += vbump;
        Kp += Kbump;
      }
      cp_async_commit();
      rd += sstride;
      if (rd == ring_end) rd = ring0;
      wr += sstride;
      if (wr == ring_end) wr = ring0;
      M::step(C, x, v, xn);
#pragma unroll
      for (int i = 0; i < X; ++i) x[i] = xn[i];
    }
    // end of observation interval k
    if (WITH_XS && a.xend_out) stcol<X>(a.xend_out + k * X * nta, nta, x);
    if (k < B.ny) {
      double cy = M::obs(x) - yk;
      if (d.noisy) {
        if (WITH_K) nk = fma(-a.sigma_lin, a.lamtot[k * NT], nk);
        cy = fma(a.sigma_y, nk, cy);
      }
      a.crow[k * NT] = cy;
    }
    if (k == B.n - 1 && B.nx > 0) {
#pragma unroll
      for (int i = 0; i < X; ++i) a.crow[(B.ny + i) * NT] = x[i] - xo[i];
    }
    if (WITH_K) {
#pragma unroll
      for (int i = 0; i < X; ++i) al[i] = aln[i];
    }
  }
  cp_async_wait<0>();
}

// in-place LU factorisation with partial pivoting of a dense n x n matrix (row-major, leading dimension
// LD), as scipy.linalg.lu_factor does for the reference's Newton solver (mici_extensions.py:745-763)
template <int LD>
MMD_D void lu_factor(double* A, int* piv, int n) {
  for (int c = 0; c < n; ++c) {
    int p = c;
    double best = fabs(A[c * LD + c]);
    for (int r = c + 1; r < n; ++r) {
      const double a = fabs(A[r * LD + c]);
      if (a > best) { best = a; p = r; }
    }
    piv[c] = p;
    if (p != c)
      for (int j = 0; j < n; ++j) { const double tv = A[c * LD + j]; A[c * LD + j] = A[p * LD + j]; A[p * LD + j] = tv; }
    const double inv = 1.0 / A[c * LD + c];
    for (int r = c + 1; r < n; ++r) {
      const double f = A[r * LD + c] * inv;
      A[r * LD + c] = f;
      for (int j = c + 1; j < n; ++j) A[r * LD + j] = fma(-f, A[c * LD + j], A[r * LD + j]);
    }
  }
}
template <int LD>
MMD_D void lu_solve(const double* A, const int* piv, int n, double* x) {
  for (int c = 0; c < n; ++c) {
    const int p = piv[c];
    if (p != c) { const double tv = x[c]; x[c] = x[p]; x[p] = tv; }
  }
  for (int i = 1; i < n; ++i) {
    double sv = x[i];
    for (int k = 0; k < i; ++k) sv = fma(-A[i * LD + k], x[k], sv);
    x[i] = sv;
  }
  for (int i = n - 1; i >= 0; --i) {
    double sv = x[i];
    for (int k = i + 1; k < n; ++k) sv = fma(-A[i * LD + k], x[k], sv);
    x[i] = sv / A[i * LD + i];
  }
}

// obs-level backward recursion: alpha_k = H_k^T lambda_k + Psib_{k+1}^T alpha_{k+1}  (J^T lambda in
// compressed form, rmult_by_jacob_constr :879-913).  Writes alpha for the block's intervals and
// returns alpha at the block start (needed for the v_0 columns of block 0).  lam: the block's rows in
// registers.  Fully unrolled over RMAX intervals so that the Psib loads (independent of the
// recursion) are issued together instead of one L2 round trip per interval.
template <class M, int NRMAX, int RMAX>
MMD_D void alpha_block(const Blk& B, const double* lam, const double* __restrict__ Psibc,
                       const double* __restrict__ xendc, int nta, double* alph_out, double* alpha_start,
                       const double* __restrict__ kapc = nullptr, double* body_bound = nullptr) {
  constexpr int X = M::X;
  // small blocks: every Psib_k of the block is fetched up front in one batch of volatile loads (one L2 round trip for
  // the whole recursion instead of one per interval; intervals beyond the block re-read its last one)
  constexpr bool PRELOAD = RMAX * X * X <= 24;
  double Pall[PRELOAD ? RMAX : 1][X * X];
  if (PRELOAD) {
#pragma unroll
    for (int k = 0; k < RMAX; ++k) {
      const int kk = k < B.n ? k : B.n - 1;
#pragma unroll
      for (int i = 0; i < X * X; ++i) Pall[PRELOAD ? k : 0][i] = ldg_vol(Psibc + (kk * X * X + i) * nta);
    }
  }
  double al[X], bnd = 0.0;
#pragma unroll
  for (int i = 0; i < X; ++i) al[i] = 0.0;
  MMD_SOLVE_UNROLL
  for (int k = RMAX - 1; k >= 0; --k) {
    if (k < B.n) {
      if (k < B.n - 1) {
        double Ps[X * X], t[X];
        if (PRELOAD) {
#pragma unroll
          for (int i = 0; i < X * X; ++i) Ps[i] = Pall[(PRELOAD && k + 1 < RMAX) ? k + 1 : 0][i];
        } else {
          ldcol_keep<X * X>(Psibc + (k + 1) * X * X * nta, nta, Ps);
        }
        mtv<X, X>(Ps, al, t);
#pragma unroll
        for (int i = 0; i < X; ++i) al[i] = t[i];
      }
      if (k < B.ny) {
        double dh[X], xe[X];
        if (!M::OBS_LINEAR) ldcol<X>(xendc + k * X * nta, nta, xe);
        M::obs_grad(xe, dh);
        const double lk = lam[k < NRMAX ? k : 0];
#pragma unroll
        for (int i = 0; i < X; ++i) al[i] = fma(dh[i], lk, al[i]);
      }
      if (k == B.n - 1 && B.nx > 0) {
        // rows ny .. ny + X - 1 hold the multipliers of the conditioned full state
#pragma unroll
        for (int r = 0; r < NRMAX; ++r)
#pragma unroll
          for (int i = 0; i < X; ++i)
            if (r == B.ny + i) al[i] += lam[r];
      }
      if (alph_out) stcol<X>(alph_out + k * X * nta, nta, al);
      if (kapc) {
        // |(J^T lam)_{v_t, j}| = |sum_i K_t[i][j] al[i]| <= sum_i kap_k[i] |al[i]| for every step t of interval k
        double kp[X], sb = 0.0;
        ldcol_keep<X>(kapc + k * X * nta, nta, kp);
#pragma unroll
        for (int i = 0; i < X; ++i) sb = fma(kp[i], fabs(al[i]), sb);
        bnd = (sb > bnd || sb != sb) ? sb : bnd;
      }
    }
  }
  {
    double Ps[X * X];
    if (PRELOAD) {
#pragma unroll
      for (int i = 0; i < X * X; ++i) Ps[i] = Pall[0][i];
    } else {
      ldcol_keep<X * X>(Psibc, nta, Ps);
    }
    mtv<X, X>(Ps, al, alpha_start);
  }
  if (body_bound) *body_bound = bnd;
}

// Woodbury solve G^{-1} r for this thread's block (lmult_by_inv_gram :915-942):
//   t_b = D_b^{-1} r_b ; s = C^{-1} sum_b A_b^T t_b ; lam_b = t_b - (D_b^{-1} A_b) s
// evaluated with the explicit block inverse and sum_b A_b^T D_b^{-1} r_b = sum_b (D_b^{-1} A_b)^T r_b (D_b symmetric):
// both products take `r` directly, so every factor load (D^-1: NRMAX (NRMAX + 1) / 2, D^-1 A: NRMAX x U values per
// thread, thread-private columns) is independent of the arithmetic and can be in flight at once -- the triangular
// solves this replaces were a chain of dependent loads and multiply-adds, and the solver calls this once per
// iteration.  `r` (registers, NRMAX entries, entries >= nrows ignored) is overwritten by lam_b; s (= u-part of
// J^T G^{-1} r) is returned in `s_out`.  `extra_max` rides along the same cross-block reduction as a maximum (the
// solver's |c|_inf).
// Loads: the factors are thread-private columns in global memory (L2 / L1 hits).  Left to the compiler, each load
// ends up next to the multiply-add that consumes it (the register cap leaves no room to hoist 57 loads), i.e. a chain
// of ~25 exposed L2 round trips per call -- 41 k cycles measured on an otherwise idle SM, once per solver iteration.
// Here they are issued in a few batches of volatile loads (kept together, in order, by the compiler), two columns of
// D^-1 plus two rows of D^-1 A per batch, and the capacitance factor is fetched before the reduction barrier: three
// round trips for the first product, two for the second.  The order of the floating-point operations is unchanged.
template <class M, int NRMAX, int UMAX, bool TAIL_SYNC = true>
MMD_D void inv_gram_block(const Dims& d, const Blk& B, bool has_blk, const double* __restrict__ Dic,
                          const double* __restrict__ DinvAc, const double* __restrict__ LCc, double* r, double* s_out,
                          double* extra_max, double* smem_red, const Tid& t, unsigned long long* ph = nullptr) {
#ifdef MMD_PHASE_CLOCK
  long long _pg = clock64();
#define PHG(i)                                                                                     \
  do {                                                                                             \
    if (threadIdx.x == 0 && ph) {                                                                  \
      const long long _n = clock64();                                                              \
      atomicAdd(&ph[i], (unsigned long long)(_n - _pg));                                           \
      _pg = _n;                                                                                    \
    }                                                                                              \
  } while (0)
#else
#define PHG(i)
#endif
  constexpr int UTRI = UMAX * (UMAX + 1) / 2;
  constexpr int CG = (NRMAX <= 8) ? 2 : 1;   // columns of D^-1 (and rows of D^-1 A) per load batch
  const int U = d.U, nta = t.nta;
  const int n = has_blk ? B.nrows : 0;
  double g[UMAX + 1], tb[NRMAX];
#pragma unroll
  for (int j = 0; j < UMAX; ++j) g[j] = 0.0;
  g[UMAX] = extra_max ? *extra_max : 0.0;
#pragma unroll
  for (int i = 0; i < NRMAX; ++i) tb[i] = 0.0;
  if (has_blk) {
    MMD_SOLVE_UNROLL
    for (int i0 = 0; i0 < NRMAX; i0 += CG) {
      if (i0 < n) {
        double dv[CG][NRMAX], av[CG][UMAX];
#pragma unroll
        for (int c = 0; c < CG; ++c) {
          const int i = i0 + c;
          if (i < NRMAX) {
            // column i of the packed symmetric inverse: entries (k, i), k >= i  (rows >= n: allocated, ignored below)
#pragma unroll
            for (int k = i; k < NRMAX; ++k) dv[c][k] = ldg_vol(Dic + tri(k, i) * nta);
#pragma unroll
            for (int j = 0; j < UMAX; ++j) av[c][j] = ldg_vol(DinvAc + (i * U + (j < U ? j : U - 1)) * nta);
          }
        }
#pragma unroll
        for (int c = 0; c < CG; ++c) {
          const int i = i0 + c;
          if (i < NRMAX && i < n) {
            const double ri = r[i];
#pragma unroll
            for (int k = i; k < NRMAX; ++k)
              if (k < n) {
                tb[k] = fma(dv[c][k], ri, tb[k]);
                if (k != i) tb[i] = fma(dv[c][k], r[k], tb[i]);
              }
#pragma unroll
            for (int j = 0; j < UMAX; ++j)
              if (j < U) g[j] = fma(av[c][j], ri, g[j]);
          }
        }
      }
    }
  }
  // factor of the capacitance matrix: fetched before the barrier of the reduction, used after it
  double LCm[UTRI];
  {
    const int utri = U * (U + 1) / 2;
#pragma unroll
    for (int i = 0; i < UTRI; ++i) {
      const double v = ldg_vol(LCc + (i < utri ? i : 0) * t.cpb);
      LCm[i] = (i < utri) ? v : 0.0;
    }
  }
  PHG(32);
  block_reduce<UMAX, 1, TAIL_SYNC>(g, smem_red, t);
  PHG(33);
  if (extra_max) *extra_max = g[UMAX];
  chol_solve_invdiag_fixed<UMAX>(LCm, U, g);
#pragma unroll
  for (int j = 0; j < UMAX; ++j) s_out[j] = g[j];
  PHG(34);
  if (has_blk) {
    constexpr int RG = (NRMAX <= 8) ? 3 : 1;   // rows of D^-1 A per load batch
    MMD_SOLVE_UNROLL
    for (int i0 = 0; i0 < NRMAX; i0 += RG) {
      if (i0 < n) {
        double av[RG][UMAX];
#pragma unroll
        for (int c = 0; c < RG; ++c)
          if (i0 + c < NRMAX) {
#pragma unroll
            for (int j = 0; j < UMAX; ++j) av[c][j] = ldg_vol(DinvAc + ((i0 + c) * U + (j < U ? j : U - 1)) * nta);
          }
#pragma unroll
        for (int c = 0; c < RG; ++c) {
          const int i = i0 + c;
          if (i < NRMAX && i < n) {
            double ti = tb[i];
#pragma unroll
            for (int j = 0; j < UMAX; ++j)
              if (j < U) ti = fma(-av[c][j], g[j], ti);
            r[i] = ti;
          }
        }
      }
    }
  }
  PHG(35);
#undef PHG
}

}  // namespace mmd
