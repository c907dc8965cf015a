"""Seeded synthetic FASTA generators for the BASELINE.json configs.

The reference ships `data_simulation/simulate_data.py`, which cannot run (it needs two git-ignored TSV files and
never seeds `random`), so the shapes are re-created here with numpy generators and fixed seeds.
All functions return ASCII `bytes` of sequence (no FASTA framing); `write_fasta` adds headers and 80-column lines.
"""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)

# simulate_data.py:11-17 uses a fixed 547-bp buffer between loci; any fixed non-repetitive sequence serves.
_BUFFER_SEQ = None


def _buffer_seq():
    global _BUFFER_SEQ
    if _BUFFER_SEQ is None:
        _BUFFER_SEQ = ACGT[np.random.default_rng(547).integers(0, 4, 547)]
    return _BUFFER_SEQ


def random_bases(rng, n):
    return ACGT[rng.integers(0, 4, n)]


def mutate(rng, unit_seq, frac, weights=(0.8, 0.1, 0.1)):
    """Apply substitutions / insertions / deletions to about `frac` of the bases."""
    seq = list(unit_seq.tolist())
    nmut = int(round(len(seq) * frac))
    if nmut == 0:
        return np.array(seq, dtype=np.uint8)
    pos = np.sort(rng.choice(len(seq), size=min(nmut, len(seq)), replace=False))[::-1]
    kinds = rng.choice(3, size=len(pos), p=np.array(weights) / sum(weights))
    for p, k in zip(pos.tolist(), kinds.tolist()):
        if k == 0:
            seq[p] = int(ACGT[(int(np.searchsorted(ACGT, seq[p])) + int(rng.integers(1, 4))) % 4])
        elif k == 1:
            seq.insert(p, int(ACGT[rng.integers(0, 4)]))
        else:
            del seq[p]
    return np.array(seq, dtype=np.uint8)


def repeat_tract(rng, m, units, frac=0.0, weights=(0.8, 0.1, 0.1)):
    motif = random_bases(rng, m)
    tract = np.tile(motif, units)
    extra = int(rng.integers(0, m))
    tract = np.concatenate([tract, motif[:extra]])
    if frac > 0:
        tract = mutate(rng, tract, frac, weights)
    return tract


def plant(rng, seq, n_repeats, m_range=(2, 100), units=(3, 30), perfect_frac=0.4, edit=(0.02, 0.12),
          weights=(0.8, 0.1, 0.1)):
    """Overwrite `n_repeats` random places of `seq` (uint8 array, modified in place) with repeat tracts."""
    L = len(seq)
    for _ in range(n_repeats):
        m = int(rng.integers(m_range[0], m_range[1] + 1))
        u = int(rng.integers(units[0], units[1] + 1))
        frac = 0.0 if rng.random() < perfect_frac else float(rng.uniform(*edit))
        t = repeat_tract(rng, m, u, frac, weights)
        if len(t) >= L:
            continue
        p = int(rng.integers(0, L - len(t)))
        seq[p:p + len(t)] = t
    return seq


def contig_c2(L=46_700_000, seed=21, n_repeats=None, n_runs=True, softmask=0.05, density_per_mbp=300):
    """chr21-scale contig: uniform background, planted perfect/impure repeats, two N runs, 5 % lower case."""
    rng = np.random.default_rng(seed)
    seq = random_bases(rng, L)
    if n_repeats is None:
        n_repeats = int(L / 1e6 * density_per_mbp)
    plant(rng, seq, n_repeats)
    if n_runs and L >= 1000:
        a = min(50_000, L // 100)
        seq[:a] = ord("N")
        b0 = int(0.4 * L)
        seq[b0:b0 + min(3_000_000, L // 15)] = ord("N")
    if softmask > 0:
        low = rng.random(L) < softmask
        seq[low] |= 0x20
    return seq.tobytes()


def contig_c1(L=10_000_000, seed=20261018):
    """simulate_data.py-style: loci separated by cyclic slices (500-3000 bp) of a fixed 547-bp buffer."""
    rng = np.random.default_rng(seed)
    buf = _buffer_seq()
    parts, total, off = [], 0, 0
    while total < L:
        n = int(rng.integers(500, 3001))
        idx = (off + np.arange(n)) % len(buf)
        off = (off + n) % len(buf)
        parts.append(buf[idx]); total += n
        m = int(rng.integers(2, 101))
        # choose_num_units (simulate_data.py:20-24): enough units for ~ 25-300 bp tracts
        u = max(3, int(rng.integers(25, 300) // m) + 2)
        frac = float(rng.uniform(0.05, 0.15))
        t = repeat_tract(rng, m, u, frac)
        parts.append(t); total += len(t)
    return np.concatenate(parts)[:L].tobytes()


def contig_c4(L=46_700_000, seed=44, n_repeats=60_000):
    rng = np.random.default_rng(seed)
    seq = random_bases(rng, L)
    plant(rng, seq, n_repeats, m_range=(30, 100), units=(3, 30), perfect_frac=0.0, edit=(0.10, 0.30),
          weights=(0.5, 0.25, 0.25))
    return seq.tobytes()


def contigs_c5(n=1_000_000, length=1000, seed=5, every=3):
    """assembly-scaffold shape: n contigs of `length` bp; a short-motif repeat in every third contig."""
    rng = np.random.default_rng(seed)
    flat = random_bases(rng, n * length).reshape(n, length)
    for i in range(0, n, every):
        m = int(rng.integers(1, 7))
        u = int(rng.integers(5, 21))
        t = np.tile(random_bases(rng, m), u)
        if rng.random() < 0.5:
            for _ in range(int(rng.integers(1, 3))):
                q = int(rng.integers(0, len(t)))
                t[q] = ACGT[rng.integers(0, 4)]
        p = int(rng.integers(0, length - len(t)))
        flat[i, p:p + len(t)] = t
    return [flat[i].tobytes() for i in range(n)]


HG38_MBP = [248, 242, 198, 190, 181, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51,
            156, 57]


def genome_c3(scale=1.0):
    """24 contigs with hg38-like lengths (sum ~3.1 Gbp at scale 1), each generated like C2 with seed 100+i."""
    return [contig_c2(int(mb * 1e6 * scale), seed=100 + i) for i, mb in enumerate(HG38_MBP)]


def fuzz_contig(rng, L, n_density=0.0, n_repeats=None, m_range=(1, 40)):
    """Small adversarial contig: repeats, isolated Ns, N runs, lower case, IUPAC codes, N at both ends."""
    seq = random_bases(rng, L)
    if L >= 20:
        if n_repeats is None:
            n_repeats = max(1, L // 150)
        plant(rng, seq, n_repeats, m_range=(m_range[0], max(m_range[0], min(m_range[1], max(1, L // 4)))),
              units=(2, 12), perfect_frac=0.5, edit=(0.03, 0.25))
    if n_density > 0 and L > 0:
        nn = rng.random(L) < n_density
        seq[nn] = ord("N")
        for _ in range(int(rng.integers(0, 3))):
            a = int(rng.integers(0, L)); b = min(L, a + int(rng.integers(1, 60)))
            seq[a:b] = ord("N")
        if rng.random() < 0.3:
            seq[:int(rng.integers(1, 4))] = ord("N")
        if rng.random() < 0.3:
            seq[L - int(rng.integers(1, min(4, L) + 1)):] = ord("n")
        if rng.random() < 0.3 and L > 10:
            seq[int(rng.integers(0, L))] = ord("R")
    if L > 0:
        low = rng.random(L) < 0.1
        seq[low] |= 0x20
    return seq.tobytes()


def write_fasta(path, contigs, names=None, width=80):
    with open(path, "wb") as f:
        for i, s in enumerate(contigs):
            name = names[i] if names else "ctg%d" % i
            f.write(b">" + name.encode() + b" synthetic\n")
            for j in range(0, len(s), width):
                f.write(s[j:j + width] + b"\n")
