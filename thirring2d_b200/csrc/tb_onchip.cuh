// tb_onchip.cuh — device helpers shared by the on-chip CG kernels (tb_resident.cu: one CTA per chain,
// tb_cluster.cu: one thread-block cluster per chain).  Include inside an anonymous namespace.
#pragma once

// o += sgn * (w * f)  and  o += sgn * (conj(w) * f), as four FMAs each (operand negation is free in SASS)
template <int SGN>
__device__ __forceinline__ void hop_acc(double2 &o, const double2 w, const double2 f) {
  if (SGN > 0) {
    o.x = fma(w.x, f.x, o.x);  o.x = fma(-w.y, f.y, o.x);
    o.y = fma(w.x, f.y, o.y);  o.y = fma(w.y, f.x, o.y);
  } else {
    o.x = fma(-w.x, f.x, o.x); o.x = fma(w.y, f.y, o.x);
    o.y = fma(-w.x, f.y, o.y); o.y = fma(-w.y, f.x, o.y);
  }
}
template <int SGN>
__device__ __forceinline__ void hopc_acc(double2 &o, const double2 w, const double2 f) {
  if (SGN > 0) {
    o.x = fma(w.x, f.x, o.x);  o.x = fma(w.y, f.y, o.x);
    o.y = fma(w.x, f.y, o.y);  o.y = fma(-w.y, f.x, o.y);
  } else {
    o.x = fma(-w.x, f.x, o.x); o.x = fma(-w.y, f.y, o.x);
    o.y = fma(-w.x, f.y, o.y); o.y = fma(w.y, f.x, o.y);
  }
}

// ---- tensor memory (TMEM) as a thread-private home for the solution vector x ------------------------------
// 256 KB per SM that nothing else on this path uses.  A warp can only reach the 32 TMEM lanes of its own quarter
// (warp % 4); within them every thread owns its lane, so "32x32b" loads/stores are exactly a per-thread scratch
// array: WORDS 32-bit columns per thread, warps that share a quarter stacked along the columns.  x += alpha p then
// costs no LSU wavefronts and no L2 round trip.
__device__ __forceinline__ void tmem_ld16(uint32_t (&v)[16], uint32_t taddr) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32"
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32"
               "[%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
               :
               : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                 "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

