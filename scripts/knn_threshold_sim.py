"""CPU simulation of the planned threshold-first candidate selection for the k-NN tensor path (DESIGN.md section 8
item 3): per query row, a threshold from a sampled pre-pass (the r-th smallest distance over every `stride`-th corpus
tile), then an append-only list of every corpus point below it.  Prints candidate counts and how often a row would fall
back to the exact kernel (fewer than k candidates, or more than the list capacity).
usage: python scripts/knn_threshold_sim.py [N] [Q]"""
import sys
import numpy as np

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 512
D, R, k, tn = 24, 4, 33, 64
rng = np.random.RandomState(5)
corpus = rng.uniform(-3.2, 3.2, (N, D))
queries = corpus[rng.choice(N, Q, replace=False)]
d2 = np.zeros((Q, N))
for r in range(R):
    s = slice(6 * r, 6 * r + 6)
    d = ((queries[:, None, s] - corpus[None, :, s]) ** 2).sum(-1)
    d2 = np.maximum(d2, d)
kth = np.partition(d2, k - 1, axis=1)[:, k - 1]
tiles = np.arange(N) // tn
print(f"N = {N}, Q = {Q}, D = {D}, {R} robots, k = {k}, corpus tiles of {tn}")
print("| sample stride | r | sampled points | mean candidates / row | p99 | rows < k (fallback) | rows > 256 | rows > 384 |")
print("|---|---|---|---|---|---|---|---|")
for stride in (6, 12, 24):
    sample = (tiles % stride) == 0
    ds = d2[:, sample]
    for r_ in (8, 12, 16):
        thr = np.partition(ds, r_ - 1, axis=1)[:, r_ - 1]
        cnt = (d2 < thr[:, None]).sum(1)
        print(f"| {stride} | {r_} | {sample.sum()} | {cnt.mean():.0f} | {np.percentile(cnt, 99):.0f} | {(cnt < k).mean() * 100:.2f} % | "
              f"{(cnt > 256).mean() * 100:.2f} % | {(cnt > 384).mean() * 100:.2f} % |")
print(f"\nstreaming heaps today: 2 halves x (k + 8) x ln(N / 2 / (k + 8)) = {2 * (k + 8) * np.log(N / 2 / (k + 8)):.0f} heap events per row")
