#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ (run in the build container, where /root/reference exists).

The reference ships no tests, seeds or golden vectors and R is not installed, so the
fixtures are: (i) the package's own datasets converted from data/*.rda (bz2-compressed RDX2 XDR
serialisations) with a small pure-Python reader, and (ii) outputs of the CPU oracle (oracle/, the line-by-line
restatement of the reference R code, pinned to 50-digit known-answer vectors) on those datasets with fixed
Philox-seeded variates.  The GPU parity tests and the oracle regression tests read only the committed .npz
files; nothing under tests/ reads /root/reference at run time.

    python tools/make_golden.py            # rewrites tests/golden/*.npz
"""
import bz2
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_DATA = "/root/reference/data"
OUT = os.path.join(ROOT, "tests", "golden")


# ---- minimal RDX2 (XDR) reader: enough for data.frames of double / integer / factor / character columns ----
class _Reader:
    def __init__(self, buf):
        self.b, self.i, self.refs = buf, 0, []

    def take(self, n):
        v = self.b[self.i:self.i + n]
        self.i += n
        return v

    def int(self):
        return struct.unpack(">i", self.take(4))[0]

    def item(self):
        flags = self.int()
        typ, has_attr, has_tag = flags & 0xFF, bool(flags & 0x200), bool(flags & 0x400)
        if typ == 254:                                  # NILVALUE
            return None
        if typ == 255:                                  # REFSXP
            return self.refs[(flags >> 8) - 1]
        if typ == 1:                                    # SYMSXP
            s = self.item()
            self.refs.append(s)
            return s
        if typ == 2:                                    # LISTSXP (pairlist)
            out = []
            while True:
                attr = self.item() if has_attr else None   # noqa: F841
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.int()
                typ, has_attr, has_tag = flags & 0xFF, bool(flags & 0x200), bool(flags & 0x400)
                if typ == 254:
                    return out
                if typ != 2:
                    raise ValueError("unexpected pairlist tail type %d" % typ)
        if typ == 9:                                    # CHARSXP
            n = self.int()
            return None if n == -1 else self.take(n).decode("latin-1")
        if typ in (10, 13):                             # LGLSXP / INTSXP
            n = self.int()
            v = np.frombuffer(self.take(4 * n), dtype=">i4").astype(np.int64)
        elif typ == 14:                                 # REALSXP
            n = self.int()
            v = np.frombuffer(self.take(8 * n), dtype=">f8").astype(np.float64)
        elif typ == 16:                                 # STRSXP
            v = [self.item() for _ in range(self.int())]
        elif typ == 19:                                 # VECSXP
            v = [self.item() for _ in range(self.int())]
        else:
            raise ValueError("unsupported SEXP type %d" % typ)
        attrs = dict(self.item()) if has_attr else {}
        return {"v": v, "attr": attrs} if attrs else v


def read_rda(path):
    buf = bz2.decompress(open(path, "rb").read())
    assert buf[:5] == b"RDX2\n" and buf[5:7] == b"X\n"
    r = _Reader(buf)
    r.i = 7
    r.int(); r.int(); r.int()                           # format version, writer R version, min reader version
    top = r.item()                                      # pairlist of (symbol, value)
    return {k: v for k, v in top}


def data_frame(obj):
    names = obj["attr"]["names"]
    names = names["v"] if isinstance(names, dict) else names
    cols = {}
    for nm, col in zip(names, obj["v"]):
        cols[nm] = col["v"] if isinstance(col, dict) else col
    return cols


def main():
    import oracle
    from oracle.gradient import gradient_draws
    from oracle.hlm import hlm_bayes_sp
    os.makedirs(OUT, exist_ok=True)

    meuse = data_frame(read_rda(os.path.join(REF_DATA, "meuse.rda"))["meuse"])
    boston = data_frame(read_rda(os.path.join(REF_DATA, "boston.c.rda"))["boston.c"])
    # config 2: meuse, coordinates in km so phi fits the default prior, y = log zinc
    m_coords = np.column_stack([np.asarray(meuse["x"], float), np.asarray(meuse["y"], float)]) / 1000.0
    m_y = np.log(np.asarray(meuse["zinc"], float))
    # config 3: boston.c, coords = (LON, LAT) * 100, y = CMEDV (vignettes/boston-vignette.Rmd:49-53)
    b_coords = np.column_stack([np.asarray(boston["LON"], float), np.asarray(boston["LAT"], float)]) * 100.0
    b_y = np.asarray(boston["CMEDV"], float)
    assert m_coords.shape == (155, 2) and b_coords.shape == (506, 2)

    def grid_over(coords, side):
        lo, hi = coords.min(0), coords.max(0)
        gx = np.linspace(lo[0] + 0.05 * (hi[0] - lo[0]), hi[0] - 0.05 * (hi[0] - lo[0]), side)
        gy = np.linspace(lo[1] + 0.05 * (hi[1] - lo[1]), hi[1] - 0.05 * (hi[1] - lo[1]), side)
        return np.array([[a, b] for b in gy for a in gx])

    def fixture(name, coords, y, cov_type, phi_range, side, S, niter, report, upper_phi):
        rng = np.random.Generator(np.random.Philox(20260101 + len(name)))
        c = 2 if cov_type == "matern1" else 5
        grid = grid_over(coords, side)
        phi = rng.uniform(*phi_range, S)
        sig2 = rng.uniform(0.5, 2.0, S)
        z = (y - y.mean())[None, :] + 0.2 * rng.standard_normal((S, y.shape[0]))
        eps = rng.standard_normal((S, grid.shape[0], c))
        draws, mean, var = gradient_draws(coords, grid, cov_type, phi, sig2, z, eps, want_moments=True)
        norm = rng.standard_normal(3 * niter)
        unif = rng.random(3 * niter)
        yc = y - y.mean()                               # the model has no mean term
        chain = hlm_bayes_sp(yc, coords, niter=niter, nburn=niter // 2, report=report, cov_type=cov_type,
                             upper_phi=upper_phi, norm=norm, unif=unif, trgt_fn_compute=True)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), coords=coords, y=yc, grid=grid, cov_type=cov_type,
                            phi=phi, sig2=sig2, z=z, eps=eps, draws=draws, mean=mean, var=var,
                            hlm_norm=norm, hlm_unif=unif, hlm_niter=niter, hlm_report=report, hlm_upper_phi=upper_phi,
                            hlm_sig2=chain["chain"]["sig2"], hlm_tau2=chain["chain"]["tau2"],
                            hlm_phi=chain["chain"]["phi"], hlm_trgt=chain["trgt_fn"], hlm_accept=chain["accept_m"],
                            hlm_tuning=chain["tuning"])
        print(name, "cond(K) max %.2e" % max(np.linalg.cond(oracle.cov_by_name(cov_type)(
            oracle.dist_matrix(coords), p, 1.0)) for p in phi))

    fixture("meuse_matern1", m_coords, m_y, "matern1", (2.0, 6.0), 8, 4, 60, 20, 30.0)
    fixture("meuse_matern2", m_coords, m_y, "matern2", (6.0, 12.0), 8, 3, 40, 20, 30.0)
    fixture("boston_matern2", b_coords, b_y, "matern2", (0.6, 1.2), 6, 2, 30, 10, 30.0)


if __name__ == "__main__":
    main()
