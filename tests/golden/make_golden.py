"""Generates the golden input/output fixtures of this directory.

    python tests/golden/make_golden.py          (run from the repo root; overwrites *.fasta / *.tsv / MANIFEST.json)

The reference is Scala on Spark + FASTdoop and cannot run in this image (no JVM), and its repository holds no test
vectors for the counting path.  The expected outputs here therefore come from `oracle/literal.py`, the
line-by-line Python transliteration of the reference's Scala (Kmer with 31-nt Long slices, getSuperKmers,
extractKXmers with the RIndex heap merge, extractKXmersHT) — the closest thing to "the reference, run here".  For
every (input, configuration) the script checks that the reference's two reduce paths agree (useHT=0 == useHT=1 as
multisets; the sort path ascending per bin, SBKC:566-597) and writes ONE file of `bin<TAB>kmer<TAB>count` lines in
bin-file order.  tests/test_golden_files.py compares the C++ oracle (CPU) and the CUDA library (GPU, both paths,
through the C ABI) with these files; it never runs this script.
"""
import hashlib
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import literal  # noqa: E402

COMP = bytes.maketrans(b"ACGT", b"TGCA")


def rnd(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def make_cases():
    rng = random.Random(20261018)
    cases = {}
    # 1. ragged reads with every kind of invalid byte (UTIL:697: anything but upper-case ACGT), reads shorter than k,
    #    an empty record, a header-only record at the end
    reads = []
    for r in range(40):
        L = rng.choice((0, 3, 19, 27, 28, 29, 40, 54, 55, 56, 80, 120))
        s = list(rnd(rng, L))
        for i in range(len(s)):
            if rng.random() < 0.03:
                s[i] = rng.choice("NnacgtRYKM-*")
        reads.append(">r%d ragged\n%s\n" % (r, "".join(s)))
    reads.append(">only a header\n")
    cases["ragged_invalid"] = "".join(reads)
    # 2. one long record on 60-column lines with a carriage return and two N runs (sequenceType=1 reads the same text)
    body = rnd(rng, 700) + "N" * 7 + rnd(rng, 500) + "NN" + rnd(rng, 391)
    lines = [body[i:i + 60] for i in range(0, len(body), 60)]
    lines[5] += "\r"
    cases["long_multiline"] = ">chrSynthetic 1600 bp\n" + "\n".join(lines) + "\n"
    # 3. homopolymers and short-period repeats: every window of a run has the same signature, counts grow large
    cases["repeats"] = (">polyA\n" + "A" * 150 + "\n>polyT\n" + "T" * 150 + "\n>polyC\n" + "C" * 90 + "\n>ac\n" + "AC" * 80 +
                        "\n>acg\n" + "ACG" * 60 + "\n>aacc\n" + "AACC" * 40 + "\n")
    # 4. both strands of the same fragments: canonical k-mers make the strands indistinguishable
    frags = [rnd(rng, rng.randint(60, 140)) for _ in range(12)]
    both = []
    for i, f in enumerate(frags):
        both.append(">f%d\n%s\n" % (i, f))
        both.append(">f%d_rc\n%s\n" % (i, f.encode().translate(COMP)[::-1].decode()))
    cases["both_strands"] = "".join(both)
    # 5. 30x coverage of a small genome with substitutions: the shape of the benchmark's reads
    genome = rnd(rng, 600)
    cov = []
    for r in range(180):
        p = rng.randint(0, len(genome) - 100)
        s = list(genome[p:p + 100])
        for i in range(100):
            if rng.random() < 0.01:
                s[i] = rng.choice("ACGT")
        s = "".join(s)
        if rng.random() < 0.5:
            s = s.encode().translate(COMP)[::-1].decode()
        cov.append(">read%d\n%s\n" % (r, s))
    cases["coverage30"] = "".join(cov)
    return cases


CONFIGS = {   # (k, m, x, max_b); b = min(4^m, max_b) (TCFG:32)
    "ragged_invalid": [(28, 10, 3, 2048), (5, 3, 1, 64), (55, 13, 3, 2048), (31, 11, 3, 4096)],
    "long_multiline": [(31, 11, 3, 4096), (33, 8, 3, 512), (64, 15, 2, 5000)],
    "repeats": [(28, 10, 3, 2048), (12, 4, 2, 100), (32, 7, 2, 333)],
    "both_strands": [(28, 10, 3, 2048), (55, 13, 3, 2048), (20, 5, 3, 2000)],
    "coverage30": [(28, 10, 3, 2048), (31, 11, 3, 4096), (60, 9, 4, 77)],
}


def main():
    manifest = []
    for name, text in make_cases().items():
        fasta = text.encode()
        with open(os.path.join(HERE, name + ".fasta"), "wb") as f:
            f.write(fasta)
        for (k, m, x, max_b) in CONFIGS[name]:
            sort_triples, per_bin = literal.count(fasta, k, m, x, max_b, 0)
            ht_triples, _ = literal.count(fasta, k, m, x, max_b, 1)
            assert sort_triples == ht_triples, (name, k, "the reference's two reduce paths disagree")
            for lines in per_bin.values():                          # bin files are ascending (SBKC:566-597)
                assert [s for s, _ in lines] == sorted(s for s, _ in lines)
            out = b"".join(b"%d\t%s\t%d\n" % (b, s.encode(), c) for b in sorted(per_bin) for s, c in per_bin[b])
            fn = "%s.k%d_m%d_x%d_b%d.tsv" % (name, k, m, x, int(min(4 ** m, max_b)))
            with open(os.path.join(HERE, fn), "wb") as f:
                f.write(out)
            manifest.append({"case": name, "fasta": name + ".fasta", "k": k, "m": m, "x": x, "max_b": max_b, "expected": fn,
                             "n_kmers": sum(c for _, _, c in sort_triples), "n_distinct": len(sort_triples),
                             "sha256": hashlib.sha256(out).hexdigest()})
            print("%-16s k=%-2d m=%-2d x=%d B=%-4d: %5d k-mers, %5d distinct, %d bytes" %
                  (name, k, m, x, max_b, manifest[-1]["n_kmers"], manifest[-1]["n_distinct"], len(out)))
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py (oracle/literal.py, the Scala transliteration)", "vectors": manifest}, f, indent=1)


if __name__ == "__main__":
    main()
