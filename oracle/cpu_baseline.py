"""CPU baseline worker: times the reference's own fmdm_invert_cg (oracle/_ref, hmc.c:341-404) or, when the
reference build is absent, the oracle port, on synthetic chains of the bench workload.  TEST/BENCH
INFRASTRUCTURE ONLY — executed only by bench.py's cpu_baseline / --impl reference legs.

    python -m oracle.cpu_baseline --nt 64 --nx 64 --chains 4 --first 0 --m 0.1 --g 0.3

prints one JSON line {"seconds", "chains", "iters", "applies", "kind"}.
"""
import argparse
import json
import time

import numpy as np

from oracle.pyoracle import MODE_ADJOINT, Oracle, RefLib, ref_available


def synthetic_chain(chain, nt, nx, g):
    """Same synthetic inputs as bench.py: quenched-equilibrium links P(A) ~ exp((Nf/g) cos A) (the fixed point
    of update_puregauge_hb, hmc.c:82-93, Nf = 2) and a complex Gaussian xi (hmc.c:439-447)."""
    rng = np.random.default_rng(1_000_003 * (chain + 1))
    A = rng.vonmises(0.0, 2.0 / g, size=(nt, nx, 2))
    xi = rng.normal(size=(nt, nx)) + 1j * rng.normal(size=(nt, nx))
    return A, xi


def family_b_inputs(nt, nx, nsrc):
    """Occupation mask (10 % monomers, fermionbag.c's field) and point sources on the first rows, as
    measure_propagator (fermionbag.c:389-435) builds them."""
    rng = np.random.default_rng(4242)
    field = (rng.random((nt, nx)) < 0.1).astype(np.int32)
    sites = [(t, x) for t in range(nt) for x in range(nx) if field[t, x] == 0][:nsrc]
    src = np.zeros((len(sites), nt, nx))
    for i, (t, x) in enumerate(sites):
        src[i, t, x] = 1.0
    return field, src


def family_b(a):
    from oracle.pyoracle import RefLibB, ref_b_available

    field, src = family_b_inputs(a.nt, a.nx, a.family_b)
    orc = Oracle()
    use_ref = ref_b_available(a.nt, a.nx)
    ref = RefLibB(a.nt, a.nx, m=a.m, mu=a.mu) if use_ref else None
    if use_ref:
        ref.set_field(field)
    t0 = time.perf_counter()
    for i in range(src.shape[0]):
        if use_ref:
            ref.call("cg_propagator", src[i])
        else:
            orc.cg_MdM(src[i], field, a.m, a.mu, propagator=True)
    secs = time.perf_counter() - t0
    print(json.dumps({"seconds": secs, "sources": int(src.shape[0]), "kind": "reference" if use_ref else "port"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nt", type=int, default=64)
    ap.add_argument("--nx", type=int, default=64)
    ap.add_argument("--chains", type=int, default=2)
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--m", type=float, default=0.1)
    ap.add_argument("--mu", type=float, default=0.0)
    ap.add_argument("--g", type=float, default=0.3)
    ap.add_argument("--max-iter", type=int, default=0)
    ap.add_argument("--flavour", default="adjoint", help="adjoint (-O3) or adjoint_shipped (the reference's own CFLAGS)")
    ap.add_argument("--family-b", type=int, default=0, metavar="NSRC",
                    help="time the reference's vec_ops.c cg_propagator on NSRC point sources instead")
    a = ap.parse_args()
    if a.family_b:
        return family_b(a)
    orc = Oracle()
    use_ref = ref_available(a.nt, a.nx, a.flavour)
    ref = RefLib(a.nt, a.nx, a.flavour, m=a.m, g=a.g, mu=a.mu) if use_ref else None
    secs, iters = 0.0, 0
    for c in range(a.first, a.first + a.chains):
        A, xi = synthetic_chain(c, a.nt, a.nx, a.g)
        if use_ref:
            G = ref.gauge(A)
            b = ref.fm_conjugate_mul(xi, G)
            t0 = time.perf_counter()
            x = ref.fmdm_invert_cg(b, G)
            secs += time.perf_counter() - t0
            # the reference does not return its iteration count; the bit-identical oracle port does
            xo, st, it, rr = orc.fmdm_invert_cg(b, A, a.m, a.mu, MODE_ADJOINT, a.max_iter)
            assert np.array_equal(x, xo)
        else:
            b = orc.fm_conjugate_mul(xi, A, a.m, a.mu, MODE_ADJOINT)
            t0 = time.perf_counter()
            xo, st, it, rr = orc.fmdm_invert_cg(b, A, a.m, a.mu, MODE_ADJOINT, a.max_iter)
            secs += time.perf_counter() - t0
        iters += it
    print(json.dumps({"seconds": secs, "chains": a.chains, "iters": iters, "applies": 2 * iters,
                      "kind": "reference" if use_ref else "port"}))


if __name__ == "__main__":
    main()
