"""Runs tests.util.oracle_accurate_window over a file of jobs in a pool of processes.

A separate interpreter (started by the test with subprocess) so that the pool forks from a process
that has never touched CUDA:   python -m tests.oracle_pool jobs.npz out.npy [procs]"""

import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests.util import oracle_accurate_window  # noqa: E402


def main():
    jobs_path, out_path = sys.argv[1], sys.argv[2]
    procs = int(sys.argv[3]) if len(sys.argv) > 3 else (os.cpu_count() or 1)
    z = np.load(jobs_path)
    fs = int(z["fs"])
    jobs = [(z["windows"][i], int(z["starts"][i]), fs, [int(b) for b in z["bits"][z["which"][i]]])
            for i in range(len(z["starts"]))]
    with mp.get_context("fork").Pool(max(1, min(procs, len(jobs)))) as pool:
        res = pool.map(oracle_accurate_window, jobs, chunksize=1)
    np.save(out_path, np.asarray(res, dtype=np.int64))


if __name__ == "__main__":
    main()
