#!/bin/bash
# tools/make_reference_goldens.sh OUTDIR — produce golden outputs of `bronko call` for BASELINE configs C1 / C2 with the
# REAL reference (treangenlab/bronko v0.1.0, Rust) and KMC3, on a machine that has both toolchains.  This image has
# neither (no cargo / rustc / kmc / kmc_tools, no network), which is why the oracle's parity for `bronko call` is labelled
# "unpinned" (DESIGN.md §6).  Run this once where the toolchains exist, commit OUTDIR/*.vcf, *.tsv, *_counts.txt.xz and
# bronko_overview.tsv under tests/golden/reference/, and tests/test_reference_goldens.py (which skips while that directory is
# absent) pins the oracle — and through it the CUDA path — on the reference's own bytes.
#
# Inputs are regenerated from the seeds below by the simulator of this repo (bronko_b200/sim.py, numpy only), so the
# goldens are reproducible on any box:  C1 = HPV16 5,000x vs the bundled hpv.bkdb;  C2 = SARS-CoV-2 (wuhan_ref) 10,000x
# vs the 4-strain k=21 db built by the reference itself.
set -euo pipefail
OUT=${1:?usage: tools/make_reference_goldens.sh OUTDIR}
REF=${BRONKO_REFERENCE:-/root/reference}          # checkout of treangenlab/bronko
REPO=$(cd "$(dirname "$0")/.." && pwd)
THREADS=${THREADS:-$(nproc)}
command -v cargo >/dev/null || { echo "cargo not found: build the reference on a box with a Rust toolchain" >&2; exit 3; }
command -v kmc >/dev/null && command -v kmc_tools >/dev/null || { echo "kmc / kmc_tools not found (bioconda: kmc)" >&2; exit 3; }
mkdir -p "$OUT"/{reads,c1,c2}

# 1. the reference binary, unmodified
( cd "$REF" && cargo build --release )
BRONKO="$REF/target/release/bronko"

# 2. the seeded inputs (bronko_b200/sim.py: SEED0 = 20251111, sample 0 of each config)
python - "$OUT/reads" <<'PY'
import sys
sys.path.insert(0, __import__("os").environ.get("REPO", "."))
from bronko_b200 import sim
out = sys.argv[1]
for cfg in ("C1", "C2"):
    r1, o1, r2, o2, truth = sim.config_reads(cfg, sample=0)
    sim.write_fastq("%s/%s_R1.fastq.gz" % (out, cfg), r1, o1, cfg, 1)
    sim.write_fastq("%s/%s_R2.fastq.gz" % (out, cfg), r2, o2, cfg, 2)
PY

# 3. C1: the bundled db;  C2: the db the reference builds from the four genomes of its own tests (tests/build_tests.rs:11-14)
"$BRONKO" call -d "$REF/test_data/hpv.bkdb" -1 "$OUT/reads/C1_R1.fastq.gz" -2 "$OUT/reads/C1_R2.fastq.gz" \
          -o "$OUT/c1" --pileup --keep-kmer-info -t "$THREADS"
G="$REPO/data/genomes"
"$BRONKO" build -g "$G/wuhan_ref.fasta" "$G/OM223929.1.fasta" "$G/ON765678.1.fasta" "$G/PX392231.1.fasta" -o "$OUT/4_sarscov2_k21" -t "$THREADS"
"$BRONKO" call -d "$OUT/4_sarscov2_k21.bkdb" -1 "$OUT/reads/C2_R1.fastq.gz" -2 "$OUT/reads/C2_R2.fastq.gz" \
          -o "$OUT/c2" --pileup --keep-kmer-info -t "$THREADS"

# 4. what to commit (the KMC dumps are large: sorted and xz-compressed; the .kmc_pre/.kmc_suf databases are not needed)
for d in c1 c2; do
    for f in "$OUT/$d"/*_counts.txt; do sort "$f" | xz -9 > "$f.xz"; rm -f "$f"; done
    rm -f "$OUT/$d"/*.kmc_pre "$OUT/$d"/*.kmc_suf
done
kmc --version 2>&1 | head -1 > "$OUT/VERSIONS.txt" || true
( cd "$REF" && git rev-parse HEAD ) >> "$OUT/VERSIONS.txt" || true
echo "goldens written to $OUT: copy c1/ c2/ VERSIONS.txt to tests/golden/reference/ and commit"
