"""Edge-case batch shared by the emulator (CPU) and GPU parity tests: QC failure, non-ACGT bases, a bad read
(nsample == 0), an over-segmented read (events/base >= 15, reference src/f5c.c:813-814), a tiny read, a read shorter
than the band, a read with exactly one k-mer more than k, and reads with out-of-range event values that must take
the exact-arithmetic instantiation of the fill kernel."""
import numpy as np

from f5c_b200 import models, synth
from f5c_b200.batch import ReadBatch


def edge_batch(model="r9", seed=9):
    b = synth.make_batch(model, n_reads=8, mean_events=400, sigma=0.3, epk=1.8, seed=seed)
    k = b.kmer_size
    seqs = [b.read_seq(i) for i in range(b.n_reads)]
    evs = [b.read_events(i).copy() for i in range(b.n_reads)]
    seqs[1] = seqs[0]                                            # events of read 1 against read 0's bases -> QC fails
    s2 = bytearray(seqs[2]); s2[10] = ord("N"); s2[50] = ord("n"); s2[-1] = ord("U"); seqs[2] = bytes(s2)
    evs[3] = np.concatenate([evs[3]] * 16)[: len(seqs[3]) * 16]   # over-segmented
    seqs[5] = seqs[5][:30]; evs[5] = evs[5][:40]                  # tiny
    seqs[6] = seqs[6][: k + 1]; evs[6] = evs[6][:5]               # two k-mers
    evs[7] = evs[7][: len(evs[7]) // 3]                           # truncated events -> cannot span / odd band path
    evs[0]["mean"][7] = 1e-30                                     # outside the fast-arithmetic range -> EXACT kernel
    evs[7]["mean"][3] = 1e-4                                      # tiny but inside it (>= 2^-60): stays FAST
    evs[2]["mean"][11] = 3e7                                      # ditto (absurd outlier event)
    good = np.ones(b.n_reads, dtype=np.uint8)
    good[4] = 0                                                   # bad read
    return ReadBatch.from_reads(seqs, evs, b.scalings.copy(), k, good=good)
