"""Seeded synthetic paired-end reads: the inputs of every BASELINE config.

150 bp paired-end reads from fixed 300 bp fragments, fragment start uniform, fragment strand
Bernoulli(0.5), i.i.d. substitution errors (rate 0.2 %, uniform over the three other bases), no
indels, no N.  10 SNVs (AF 1.0) and 20 iSNVs (AF in {0.03, 0.05, 0.10, 0.20, 0.40}, four each) are
planted on the source genome, >= 100 bp from the ends and >= 30 bp apart.  numpy only.
"""
import gzip
import os

import numpy as np

SEED0 = 20251111
GENOME_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "genomes")
SARS4 = ["wuhan_ref.fasta", "OM223929.1.fasta", "ON765678.1.fasta", "PX392231.1.fasta"]  # reference tests/build_tests.rs:11-14
HPV16 = "HPV16.fa"

_CODE = np.full(256, 0, dtype=np.uint8)
for _i, _c in enumerate("ACGT"):
    _CODE[ord(_c)] = _i
    _CODE[ord(_c.lower())] = _i
_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)


def genome_path(name):
    return os.path.join(GENOME_DIR, name)


def read_fasta(path):
    """[(header, bases as uint8 ASCII array)]"""
    opener = gzip.open if path.endswith(".gz") else open
    recs = []
    with opener(path, "rb") as f:
        name, chunks = None, []
        for line in f:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                if name is not None:
                    recs.append((name, np.frombuffer(b"".join(chunks), dtype=np.uint8)))
                name, chunks = line[1:].decode(), []
            elif name is not None:
                chunks.append(line)
        if name is not None:
            recs.append((name, np.frombuffer(b"".join(chunks), dtype=np.uint8)))
    return recs


def write_fasta(path, name, codes):
    seq = _ASCII[codes].tobytes().decode()
    with open(path, "w") as f:
        f.write(">%s\n" % name)
        for i in range(0, len(seq), 60):
            f.write(seq[i:i + 60] + "\n")


def plant_variants(L, rng, n_snv=10, n_isnv=20, margin=100, spacing=30):
    """Positions + AFs of the planted variants (drawn once per config from the seed)."""
    pos = []
    while len(pos) < n_snv + n_isnv:
        p = int(rng.integers(margin, L - margin))
        if all(abs(p - q) >= spacing for q in pos):
            pos.append(p)
    afs = [1.0] * n_snv + [af for af in (0.03, 0.05, 0.10, 0.20, 0.40) for _ in range(n_isnv // 5)]
    return np.array(pos[:n_snv + n_isnv]), np.array(afs[:n_snv + n_isnv])


def mutate_genome(codes, rate, seed):
    """i.i.d. substitutions at `rate` (the synthetic strains of config C5)."""
    rng = np.random.default_rng(seed)
    out = codes.copy()
    hit = rng.random(len(codes)) < rate
    out[hit] = (out[hit] + rng.integers(1, 4, size=int(hit.sum()), dtype=np.uint8)) & 3
    return out


def simulate_pairs(genome_ascii, depth, seed, read_len=150, frag_len=300, err=0.002, n_snv=10, n_isnv=20,
                   chunk_pairs=200_000):
    """Returns (r1 bases, r1 offsets, r2 bases, r2 offsets, truth) — bases are ASCII uint8 arrays
    (concatenated reads), offsets uint32 arrays with n_pairs+1 entries."""
    rng = np.random.default_rng(seed)
    g = _CODE[np.asarray(genome_ascii, dtype=np.uint8)]
    L = len(g)
    pos, afs = plant_variants(L, rng, n_snv, n_isnv)
    alt = (g[pos] + rng.integers(1, 4, size=len(pos), dtype=np.uint8)) & 3
    h0 = g.copy()
    fixed = afs >= 1.0
    h0[pos[fixed]] = alt[fixed]
    n_pairs = int(round(depth * L / frag_len))
    r1 = np.empty((n_pairs, read_len), dtype=np.uint8)
    r2 = np.empty((n_pairs, read_len), dtype=np.uint8)
    ar = np.arange(frag_len)
    for c0 in range(0, n_pairs, chunk_pairs):
        n = min(chunk_pairs, n_pairs - c0)
        start = rng.integers(0, L - frag_len + 1, size=n)
        frag = h0[start[:, None] + ar[None, :]]
        for p, a, af in zip(pos[~fixed], alt[~fixed], afs[~fixed]):
            cover = (start <= p) & (p < start + frag_len)
            carry = cover & (rng.random(n) < af)
            frag[carry, p - start[carry]] = a
        strand = rng.random(n) < 0.5
        frag[strand] = 3 - frag[strand][:, ::-1]
        a1 = frag[:, :read_len].copy()
        a2 = (3 - frag[:, ::-1])[:, :read_len].copy()
        for arr in (a1, a2):
            e = rng.random(arr.shape, dtype=np.float32) < err
            arr[e] = (arr[e] + rng.integers(1, 4, size=int(e.sum()), dtype=np.uint8)) & 3
        r1[c0:c0 + n] = _ASCII[a1]
        r2[c0:c0 + n] = _ASCII[a2]
    off = (np.arange(n_pairs + 1, dtype=np.uint64) * read_len).astype(np.uint32)
    truth = {"pos": pos, "ref": g[pos], "alt": alt, "af": afs}
    return r1.reshape(-1), off, r2.reshape(-1), off.copy(), truth


def write_fastq(path, bases, off, tag="s", mate=1):
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "wb") as f:
        raw = bases.tobytes()
        for i in range(len(off) - 1):
            s = raw[off[i]:off[i + 1]]
            f.write(b"@%s_%d/%d\n%s\n+\n%s\n" % (tag.encode(), i, mate, s, b"I" * len(s)))


def load_genome(name):
    return read_fasta(genome_path(name))[0][1]


def config_reads(config, sample=0, depth=None):
    """Reads of one sample of a BASELINE config: 'C1' HPV16 5,000x, 'C2' SARS-CoV-2 10,000x (wuhan),
    'C4' sample s from strain s mod 4 at 10,000x."""
    if config == "C1":
        g, d = load_genome(HPV16), 5000
    elif config == "C2":
        g, d = load_genome(SARS4[0]), 10000
    elif config == "C4":
        g, d = load_genome(SARS4[sample % 4]), 10000
    else:
        raise ValueError(config)
    return simulate_pairs(g, depth if depth is not None else d, SEED0 + sample)


# ---- the same recipe on the GPU (torch), for inputs too large to simulate on the host -------------------------------
def plant_for(genome_ascii, seed, n_snv=10, n_isnv=20):
    """The planted variants of a sample (drawn once from the sample's seed, as simulate_pairs does): every chunk /
    rank of a deep sample must carry the same ones."""
    rng = np.random.default_rng(seed)
    g = _CODE[np.asarray(genome_ascii, dtype=np.uint8)]
    pos, afs = plant_variants(len(g), rng, n_snv, n_isnv)
    alt = (g[pos] + rng.integers(1, 4, size=len(pos), dtype=np.uint8)) & 3
    return pos, alt, afs


def simulate_pairs_torch(genome_ascii, n_pairs, seed, device, plan, read_len=150, frag_len=300, err=0.002, batch=400_000):
    """n_pairs read pairs of one chunk of a sample, generated on `device` with torch's Philox generator seeded by
    `seed` (reproducible on the same GPU type and torch version: bench.py's sharded leg regenerates the chunks of
    other ranks from their seeds instead of moving them).  Returns (r1 bases, r2 bases, offsets): uint8 device tensors of
    n_pairs * read_len + 64 bytes (64 bytes of '*' slack) and one int32 offsets tensor (n_pairs + 1) valid for both."""
    import torch
    pos, alt, afs = plan
    g = torch.from_numpy(_CODE[np.asarray(genome_ascii, dtype=np.uint8)]).to(device)
    L = g.numel()
    fixed = afs >= 1.0
    h0 = g.clone()
    h0[torch.from_numpy(pos[fixed]).to(device)] = torch.from_numpy(alt[fixed]).to(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    lut = torch.from_numpy(_ASCII.copy()).to(device)
    out1 = torch.full((n_pairs * read_len + 64,), ord("*"), dtype=torch.uint8, device=device)
    out2 = torch.full((n_pairs * read_len + 64,), ord("*"), dtype=torch.uint8, device=device)
    ar = torch.arange(frag_len, device=device, dtype=torch.int32)
    for c0 in range(0, n_pairs, batch):
        n = min(batch, n_pairs - c0)
        start = torch.randint(0, L - frag_len + 1, (n,), generator=gen, device=device, dtype=torch.int32)
        frag = h0[(start[:, None] + ar[None, :]).long()]
        for p, a, af in zip(pos[~fixed].tolist(), alt[~fixed].tolist(), afs[~fixed].tolist()):
            carry = (start <= p) & (start + frag_len > p) & (torch.rand(n, generator=gen, device=device) < af)
            rows = carry.nonzero(as_tuple=True)[0]
            frag[rows, (p - start[rows]).long()] = a
        strand = torch.rand(n, generator=gen, device=device) < 0.5
        frag = torch.where(strand[:, None], 3 - frag.flip(1), frag)
        for out, a in ((out1, frag[:, :read_len]), (out2, (3 - frag.flip(1))[:, :read_len])):
            e = torch.rand((n, read_len), generator=gen, device=device) < err
            sub = torch.randint(1, 4, (n, read_len), generator=gen, device=device, dtype=torch.uint8)
            a = torch.where(e, (a + sub) & 3, a)
            out[c0 * read_len:(c0 + n) * read_len] = lut[a.long()].reshape(-1)
    off = (torch.arange(n_pairs + 1, device=device, dtype=torch.int64) * read_len).to(torch.int32)
    return out1, out2, off
