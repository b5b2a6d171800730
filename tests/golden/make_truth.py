"""Extended-precision yardstick values for a sample of the D(omega) fixture points (tests/golden/disp.json).

For each chosen point: D in 40-digit arithmetic (tests/truth_mp.py) from the SAME inputs the reference's hot path
sees -- rho_d, J0, J1 as the reference computed them on the host (background.json) -- so that the only differences
left are the quadrature of the seven M integrals and the determinant.  Needs neither the reference nor a GPU:
    python tests/golden/make_truth.py        -> tests/golden/truth.json
"""
import json
import os
import sys
import multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import truth_mp  # noqa: E402

# (case index in disp.json, point indices)
PICK = [(0, [0, 3]), (1, [0, 5, 9]), (2, [0, 4]), (3, [1, 6]), (4, [0, 6, 11]), (5, [0, 4]), (6, [1, 3]),
        (7, [0, 3]), (8, [1]), (9, [0, 4, 8]), (10, [0, 2])]
# (PowerBumpTail is left out: sigma has a kink at aBT that is not among SizeDensity.poles, so the reference's own
# quadrature -- which the product follows node for node -- carries a 1e-9 error there that an exact integral lacks)


def one(job):
    ci, pi = job
    disp = json.load(open(os.path.join(HERE, "disp.json")))["cases"][ci]
    bg = json.load(open(os.path.join(HERE, "background.json")))[disp["dist"]]
    spec = bg["spec"]
    denom = bg["vgy"]**2 + bg["vgx"]**2/4
    j01 = (-bg["vgy"]/denom - 1.0, bg["vgx"]/(2*denom))
    if spec["kind"] == "powerlaw":
        t = truth_mp.TruthDispersion(spec["mu"], spec["stokes_range"], spec.get("c", 20.0), spec.get("power", 3.5),
                                     rhod=bg["rhod"], j01=j01)
    else:
        t = truth_mp.TruthDispersion(spec["mu"], [spec["amin"], spec["aR"]], spec.get("c", 20.0),
                                     bump={k: spec.get(k) for k in ("amin", "aP", "aL", "aR", "bumpfac", "aBT")},
                                     poles=bg["poles"], rhod=bg["rhod"], j01=j01)
    w = complex(*disp["w"][pi])
    D, M = t.calculate(w, disp["kx"], disp["kz"], disp["alpha"], return_M=True)
    # natural scale of the determinant: sum of the magnitudes of its expansion terms (a root of D is a point
    # where they cancel; errors are measured against this scale, not against |D|)
    import numpy as np
    sys.path.insert(0, os.path.dirname(HERE))
    from common import det_terms_scale
    scale = det_terms_scale(M, w, disp["kx"], disp["kz"], disp["alpha"], spec.get("c", 20.0), bg["vgx"])
    return dict(case=ci, point=pi, scale=scale, dist=disp["dist"], kx=disp["kx"], kz=disp["kz"], alpha=disp["alpha"],
                w=[w.real, w.imag], D_truth=[D.real, D.imag], M_truth=[[m.real, m.imag] for m in M],
                D_ref=disp["D"][pi], quad_rel_err=t.max_err)


if __name__ == "__main__":
    jobs = [(ci, pi) for ci, pts in PICK for pi in pts]
    with mp.get_context("fork").Pool(min(8, os.cpu_count() or 1)) as pool:
        res = pool.map(one, jobs, chunksize=1)
    for r in res:
        dt, dr = complex(*r["D_truth"]), complex(*r["D_ref"])
        print("%-9s K=(%g,%g) a=%g w=%s  |ref-truth|/scale = %.2e  quad err %.1e"
              % (r["dist"], r["kx"], r["kz"], r["alpha"], complex(*r["w"]), abs(dr - dt)/r["scale"], r["quad_rel_err"]))
    with open(os.path.join(HERE, "truth.json"), "w") as fh:
        json.dump(dict(digits=40, points=res), fh, indent=0)
