"""Native featuriser: thread scaling of one sss_featurize_batches call on this host:
python scripts/r2_featurize_threads.py [n_sessions]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sessionsimilaritysearch_b200 import featurize, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    sess = synth.make_sessions(n, 5)
    flat = featurize.flatten(sess, featurize.QueryVocab())
    for th in (1, 2, 4, 8, 1, 8):
        t0 = time.perf_counter()
        for _ in range(3):
            featurize.featurize_arrays(flat, n_threads=th)
        print("%d sessions, %d threads: %.2f ms per call" % (n, th, (time.perf_counter() - t0) / 3 * 1e3))


if __name__ == "__main__":
    main()
