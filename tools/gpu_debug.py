"""Developer diagnostics for a GPU box: prints parity numbers instead of asserting."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ba_oracle  # noqa: E402
from sat_bundleadjust_b200 import ba_core, synth  # noqa: E402
from sat_bundleadjust_b200.solver import DeviceProblem, initial_vars  # noqa: E402
import util  # noqa: E402


def main():
    G = util.load_ba_golden()
    for name in [str(s) for s in G["cases"]]:
        pre = name + "/"
        p = util.params_from_golden(G, name)
        try:
            prob = DeviceProblem(p)
        except NotImplementedError as e:
            print(name, "-> not implemented:", e)
            continue
        x0 = initial_vars(p)
        r, cost = prob.residuals(x0)
        ref = G[pre + "ref_fun_x0"]
        print("%-18s fun max|diff| %.3e  (|r|max %.3e) cost %.6e vs %.6e" % (
            name, np.abs(r - ref).max(), np.abs(ref).max(), cost, 0.5 * ref @ ref))
        if p.pts_ind.size <= 1300 and "RTK" not in name:
            Jc, Jp = prob.jacobian_blocks(x0)
            J = util.dense_jacobian_from_blocks(p, Jc, Jp)
            Jfd = ba_oracle.dense_jacobian_fd(x0.copy(), p, rel_step=1e-7)
            scale = np.abs(Jfd).max(axis=0) + 1e-30
            print("   J vs FD: max col-relative diff %.3e" % (np.abs(J - Jfd) / scale).max())
            U, V, g = prob.normal_blocks(x0, "soft_l1", 1.0)
            # robust-rescaled dense reference from the GPU's own J
            f = ba_oracle.residuals(x0.copy(), p)
            z = f ** 2
            rho1, rho2 = (1 + z) ** -0.5, -0.5 * (1 + z) ** -1.5
            js = np.sqrt(np.maximum(rho1 + 2 * rho2 * f ** 2, 2.2e-16))
            Js = J * js[:, None]
            fs = f * rho1 / js
            H = Js.T @ Js
            gd = Js.T @ fs
            c = p.n_params
            Ud = np.array([H[j * c:(j + 1) * c, j * c:(j + 1) * c] for j in range(p.n_cam)])
            off = p.n_cam * c
            Vd = np.array([[H[off + 3 * i + a, off + 3 * i + b] for a, b in ((0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2))]
                           for i in range(p.n_pts)])
            print("   U rel %.3e  V rel %.3e  g rel %.3e" % (
                np.abs(U - Ud).max() / np.abs(Ud).max(), np.abs(V - Vd).max() / np.abs(Vd).max(),
                np.abs(g - gd).max() / np.abs(gd).max()))
        if pre + "ref_vars_ba" in G:
            cfg = util.ls_from_golden(G, name)
            t = time.time()
            x, rr, info = prob.solve(x0, loss=cfg.get("loss", "linear"), f_scale=cfg.get("f_scale", 1.0),
                                     max_nfev=cfg.get("max_iter", 300), verbose=0)
            dt = time.time() - t
            loss, fsc = cfg.get("loss", "linear"), cfg.get("f_scale", 1.0)
            c_ref = ba_oracle.robust_cost(ba_oracle.residuals(G[pre + "ref_vars_ba"].copy(), p), loss, fsc)
            c_tight = ba_oracle.robust_cost(G[pre + "ref_tight_fun"], loss, fsc)
            c_gpu = ba_oracle.robust_cost(ba_oracle.residuals(x.copy(), p), loss, fsc)
            print("   solve default: gpu cost %.9e (info %.9e) nfev %d it %d status %d retries %d | ref %.9e nfev %d | ref tight %.9e | %.1f ms wall, %.2f ms dev" % (
                c_gpu, info["cost"], info["nfev"], info["iterations"], info["status"], info["chol_retries"], c_ref,
                int(G[pre + "ref_nfev"]), c_tight, dt * 1e3, info["solve_ms"]))
            x, rr, info = prob.solve(x0, loss=loss, f_scale=fsc, max_nfev=400, ftol=1e-14, xtol=1e-14, gtol=1e-14)
            c_gpu_t = ba_oracle.robust_cost(ba_oracle.residuals(x.copy(), p), loss, fsc)
            print("   solve tight:   gpu cost %.12e nfev %d it %d status %d | ref tight %.12e  rel diff %.3e" % (
                c_gpu_t, info["nfev"], info["iterations"], info["status"], c_tight, (c_gpu_t - c_tight) / c_tight))
        prob.close()

    # mid-size run
    for model, corr in (("perspective", ["R", "T"]), ("affine", ["R", "T"])):
        sc = synth.make_scene(n_cam=10, n_tracks=20000, p_vis=0.5, cam_model=model, seed=1)
        p = synth.scene_to_params(sc, corr)
        ls = {"loss": "soft_l1", "f_scale": 1.0, "max_iter": 300, "verbose": 0}
        t = time.time()
        out = ba_core.run_ba_optimization(p, ls, False, False, return_info=True)
        dt = time.time() - t
        info = out[-1]
        print(model, "K=%d" % p.n_obs, "gpu: cost %.9e nfev %d it %d status %d dev %.2f ms wall %.1f ms launches %d; err %.3f -> %.3f" % (
            info["cost"], info["nfev"], info["iterations"], info["status"], info["solve_ms"], dt * 1e3, info["gpu_launches"],
            out[2].mean(), out[3].mean()))
        t = time.time()
        o = ba_oracle.solve(p, ls)
        print("   oracle: cost %.9e nfev %d  %.1f s; err -> %.3f" % (
            ba_oracle.robust_cost(ba_oracle.residuals(o[1].copy(), p), "soft_l1", 1.0), o[4], time.time() - t, o[3].mean()))


if __name__ == "__main__":
    main()
