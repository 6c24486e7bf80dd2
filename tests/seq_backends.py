"""CPU-oracle back end of the two-stage sequence driver (h-slam_b200/sequence_ba.py).  TEST INFRASTRUCTURE: it runs the
same host loop as the CUDA back end with every hot-path call answered by the oracle, so that trajectories, residual states
and thresholds of the two can be compared."""
import numpy as np


class _OrcTracker:
    def __init__(self, O, cam, levels, pc, ref_aff, prec):
        self.O, self.levels, self.prec = O, levels, prec
        self.orc = O.TrackerOracle(cam, levels, pc, ref_aff=ref_aff)

    def track(self, img, T0, aff0, coarsest):
        _, _, dI = self.O.pyramid(img, self.levels)
        r = self.orc.track(dI, T0, aff0, coarsest, prec=self.prec)
        return r["ok"], r["T"], r["aff"]

    def close(self):
        pass


class _OrcBA:
    def __init__(self, O, win, handles, prec, threads=1):
        self.win = win
        self.ba = O.BAOracle(win, [h["dI0"] for h in handles], prec=prec, threads=threads, fast=threads > 1)

    def set_tables(self, win):
        self.ba.set_window(win)

    def update_points(self, idepth, deltaF, idepth_zero):
        self.ba.update_points(idepth=idepth, deltaF=deltaF, idepth_zero=idepth_zero)

    def linearize(self, fix):
        return self.ba.linearize(fix)

    def new_frame_energy_th(self, newest):
        return self.ba.new_frame_energy_th(newest)

    def apply(self):
        self.ba.apply_res()

    def accumulate(self):
        return [self.ba.accumulate(k)[:2] for k in range(3)]

    def resubstitute(self, x):
        return self.ba.resubstitute(x)

    def close(self):
        pass


class OracleBackend:
    def __init__(self, O, prec="f", threads=1):
        self.O, self.prec, self.threads = O, prec, threads

    def new_frame(self, img, levels):
        lv, _, dI_flat = self.O.pyramid(img, levels)
        return dict(dI0=lv[0], dI_flat=dI_flat)

    def free_frame(self, handle):
        pass

    def open_ba(self, win, handles):
        return _OrcBA(self.O, win, handles, self.prec, self.threads)

    def _cloud(self, cam, levels, kf, cpt, hdi):
        su = (cpt[:, 0].astype(np.float32) + np.float32(0.5)).astype(np.int32)
        sv = (cpt[:, 1].astype(np.float32) + np.float32(0.5)).astype(np.int32)
        sw = np.sqrt((1e-3 / (hdi.astype(np.float64) + 1e-12)).astype(np.float32)).astype(np.float32)   # CoarseTracker.cpp:122
        return self.O.make_coarse_depth(levels, cam.w, cam.h, su, sv, cpt[:, 2].astype(np.float32), sw, kf.handle["dI_flat"])

    def tracker_from_points(self, cam, levels, kf):
        cpt = np.stack([kf.u, kf.v, kf.idepth], 1).astype(np.float32)
        pc = self._cloud(cam, levels, kf, cpt, np.full(len(kf.u), 1e-3, np.float32))
        return _OrcTracker(self.O, cam, levels, pc, kf.aff_now(), self.prec)

    def tracker_from_ba(self, cam, levels, kf, ba, win, lin, ids):
        cpt = np.asarray(lin["centerProjectedTo"], np.float32)[ids]
        hdi = np.asarray(ba.ba.point_results()["HdiF"], np.float32)[win.point_idx[ids]]
        pc = self._cloud(cam, levels, kf, cpt, hdi)
        return _OrcTracker(self.O, cam, levels, pc, kf.aff_now(), self.prec)
