import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from oracle import synth, glue, replay
from ros_stereo_slam_b200 import VisualFrontEnd
np.set_printoptions(precision=17, linewidth=200)
fe = VisualFrontEnd(ransac_exhaustive=1)
for n, frac in ((500, 0.1), (5000, 0.3)):
    X, xy, _, _, _ = synth.pnp_stress_case(n, frac, 0.3, seed=3)
    r0 = replay.pnp_ransac(X, xy, glue.K, 100, 1.0, 0.99, exhaustive=True)
    r = fe.solvePnPRansac(X, xy, 100, 1.0, 0.99)
    last = fe.last_pnp()
    hyp = np.array([np.concatenate(h) for h in r0["hyp"]])
    d = np.abs(last["models"] - hyp).max(1)
    print("n", n, "hyp maxdiff", d.max(), "n exact", (d == 0).sum(), "of", len(d))
    print(" worst idx", np.argsort(-d)[:5], np.sort(-d)[:5])
    print(" counts equal", np.array_equal(last["counts"], r0["counts"]), "ndiff", (last["counts"] != r0["counts"]).sum())
    bad = np.nonzero(last["counts"] != r0["counts"])[0][:10]
    print(" bad", bad, last["counts"][bad], r0["counts"][bad], d[bad])
    print(" best", last["best"], r0["best"], "niters", last["n_iters"], r0["n_iters"])
    print(" inliers equal", np.array_equal(r["inliers"], r0["inliers"]), len(r["inliers"]), len(r0["inliers"]))
    print(" pose diff", np.abs(r["rvec"] - r0["rvec"]).max(), np.abs(r["tvec"] - r0["tvec"]).max())
    i = int(np.argmax(d))
    print(" gpu", last["models"][i]); print(" cv2", hyp[i])
