"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference).

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference cannot travel to
the GPU box):

    python oracle/gen_golden.py

Every fixture stores its inputs next to the reference's outputs so the parity tests do
not depend on RNG reproducibility.  Sizes are kept to a few hundred KB in total.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


class _Src:
    """Minimal stand-in for source.IQwav: .length only (chunker.py:32 reads .length)."""

    def __init__(self, n):
        self.length = n


def noise_c64(rng, n, scale=40.0):
    return ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * scale).astype(np.complex64)


def fm_tone_c64(rng, n, fs, f_off, f_mod, beta, amp=60.0, noise=2.0):
    t = np.arange(n) / fs
    ph = 2 * np.pi * f_off * t + beta * np.sin(2 * np.pi * f_mod * t)
    x = amp * np.exp(1j * ph) + noise * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return x.astype(np.complex64)


def main():
    dd = ref_shim.load()
    from directdemod import chunker, comm, constants, demod_am, demod_fm, filters

    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20260101)

    # ---- 1. headline chain: offsetFreq -> bh151 -> bwLim -> demod_fm, chunked ------
    fs = 2048000
    for name, x, f_off, bw in (
        ("chain_noise_d34", noise_c64(rng, 9000), 30000, 60000),
        ("chain_fmtone_d34", fm_tone_c64(rng, 9000, fs, 30000, 2400, 3.0), 30000, 60000),
        ("chain_fmtone_d68", fm_tone_c64(rng, 9000, fs, -12500.5, 1000, 2.0), -12500.5, 30000),
        ("chain_noise_d50", noise_c64(rng, 9000), 1234.5, 40960),   # 2048000/40960 = 50
    ):
        rec = {"x": x, "fs": fs, "f_off": f_off, "bw": bw}
        for tag, csize in (("whole", len(x) + 1), ("c2500", 2500), ("c1111", 1111), ("c97", 97)):
            ck = chunker.chunker(_Src(len(x)), csize)
            bh = filters.blackmanHarris(151)
            fm = demod_fm.demod_fm()
            out = comm.commSignal(1)
            outc = []
            for a, b in ck.getChunks:
                s = comm.commSignal(fs, x[a:b], ck).offsetFreq(f_off).filter(bh).bwLim(bw, uniq="First")
                outc.append(np.array(s.signal))
                s.funcApply(fm.demod)
                out.extend(s)
            rec["fm_" + tag] = out.signal
            rec["iq_" + tag] = np.concatenate(outc)
            rec["rate"] = out.sampRate
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)

    # ---- 2. mixer alone, huge global index (1-hour 2.4 Msps stream end) -----------
    x = noise_c64(rng, 4096)
    rec = {"x": x}
    for tag, fsx, f, n0 in (("a", 2400000, 100000.0, 8639990000), ("b", 2048000, 30000, 0),
                            ("c", 2048000, -777.25, 1843100000)):
        ck = chunker.chunker(_Src(10), 10)
        ck.set(constants.CHUNK_FREQOFFSET, n0)
        s = comm.commSignal(fsx, x, ck).offsetFreq(f)
        rec["y_" + tag] = s.signal
        rec["p_" + tag] = np.array([fsx, f, n0], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "mixer.npz"), **rec)

    # ---- 3. stateful FIR / IIR over chunks, complex and real inputs ----------------
    xc = noise_c64(rng, 3000)
    xr = rng.standard_normal(3000).astype(np.float32)
    cuts = [0, 700, 701, 2000, 3000]
    rec = {"xc": xc, "xr": xr, "cuts": np.array(cuts)}
    mk = {
        "bh151": lambda: filters.blackmanHarris(151),
        "ham492": lambda: filters.hamming(492),
        "gauss51": lambda: filters.gaussian(51, 5),
        "roll7": lambda: filters.rollingAverage(7),
        "remez255": lambda: filters.remez(2400000, [[0, 100000], [120000, 1199999]], [1, 0], ntaps=255),
        "butlp8": lambda: filters.butter(2400000, 100000, n=8),
        "butbp6": lambda: filters.butter(60235, 400, 4400, n=6, typeFlt=constants.FLT_BP),
        "buthp3": lambda: filters.butter(48000, 3000, n=3, typeFlt=constants.FLT_HP),
    }
    for k, ctor in mk.items():
        for tag, x in (("c", xc), ("r", xr)):
            f = ctor()
            rec["%s_%s" % (k, tag)] = np.concatenate(
                [f.applyOn(x[cuts[i]:cuts[i + 1]]) for i in range(len(cuts) - 1)])
        f = ctor()
        rec[k + "_b"] = np.asarray(f.getB, dtype=np.float64)
        rec[k + "_a"] = np.asarray(f.getA, dtype=np.float64)
    # stateless and zero-phase variants
    for k, kw in (("bh151", dict(n=151)), ("ham492", dict(n=492))):
        cls = filters.blackmanHarris if k == "bh151" else filters.hamming
        rec[k + "_zp_r"] = cls(kw["n"], zeroPhase=True).applyOn(xr)
        rec[k + "_zp_c"] = cls(kw["n"], zeroPhase=True).applyOn(xc)
        rec[k + "_sl_r"] = cls(kw["n"], storeState=False).applyOn(xr)
    rec["butbp6_zp_r"] = filters.butter(60235, 400, 4400, n=6, typeFlt=constants.FLT_BP,
                                        zeroPhase=True).applyOn(xr)
    rec["butlp8_sl_c"] = filters.butter(2400000, 100000, n=8, storeState=False).applyOn(xc)
    np.savez_compressed(os.path.join(OUT, "filters.npz"), **rec)

    # ---- 4. FM / AM demodulators -------------------------------------------------
    x = fm_tone_c64(rng, 5000, 60235, 0.0, 2400, 1.5, amp=30.0, noise=1.0)
    rec = {"x": x, "cuts": np.array([0, 1, 2, 700, 5000])}
    cuts = [0, 1, 2, 700, 5000]
    fm = demod_fm.demod_fm()
    rec["fm_state"] = np.concatenate([fm.demod(x[cuts[i]:cuts[i + 1]]) for i in range(4)])
    rec["fm_whole"] = demod_fm.demod_fm(storeState=False).demod(x)
    cuts2 = [0, 3, 700, 5000]
    fmad = demod_fm.demod_fmAD()
    rec["fmad_state"] = np.concatenate([fmad.demod(x[cuts2[i]:cuts2[i + 1]]) for i in range(3)])
    am = demod_am.demod_am()
    a = (1.0 + 0.5 * np.sin(2 * np.pi * 30 * np.arange(2400) / 2400)) * np.sin(
        2 * np.pi * 300 * np.arange(2400) / 2400) + 0.01 * rng.standard_normal(2400)
    a = a.astype(np.float32)
    rec["am_x"] = a
    rec["am_even"] = am.demod(a)
    rec["am_odd"] = am.demod(a[:2187])
    rec["am_small"] = am.demod(a[:30])
    amf = demod_am.demod_amFLT(20800, 1200)
    rec["amflt"] = np.concatenate([amf.demod(a[:1000]), amf.demod(a[1000:])])
    np.savez_compressed(os.path.join(OUT, "demod.npz"), **rec)

    # ---- 5. bwLim: integer decimation with chunker carry + strict resample ---------
    x = rng.standard_normal(5883).astype(np.float32)
    rec = {"x": x}
    ck = chunker.chunker(_Src(len(x)), 1000)
    parts = []
    for a0, b0 in ck.getChunks:
        parts.append(comm.commSignal(2048000, x[a0:b0], ck).bwLim(60000, uniq="q").signal)
    rec["dec34"] = np.concatenate(parts)
    s = comm.commSignal(60235, x).bwLim(20800, True)
    rec["strict_even"] = s.signal            # 5883 -> 2031
    rec["strict_rate"] = np.array([s.sampRate])
    rec["strict_b"] = comm.commSignal(60235, x[:5800]).bwLim(40960, True).signal   # -> 3944 (even)
    rec["strict_c"] = comm.commSignal(48000, x[:4801]).bwLim(12000, True).signal   # -> 1200
    np.savez_compressed(os.path.join(OUT, "bwlim.npz"), **rec)

    # ---- 6. chunker bounds -------------------------------------------------------
    rec = {}
    for ln, sz in ((0, 10), (5, 10), (10, 10), (20, 10), (25, 10), (100, 7), (1, 1), (93 * 20000000 - 1, 20000000)):
        rec["L%d_S%d" % (ln, sz)] = np.array(chunker.chunker(_Src(ln), sz).getChunks, dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "chunker.npz"), **rec)

    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print("  %-28s %8d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))


if __name__ == "__main__":
    main()
