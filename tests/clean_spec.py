"""Clean specification of the observable semantics, ~20 lines of
pure Python, used to cross-check the oracle on small inputs.  TEST INFRASTRUCTURE."""
from collections import Counter

COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def hash_to_bucket(s, B):                                  # UTIL:686-695, 32-bit wrapping
    M = 0xFFFFFFFF
    key = s & M
    key = ((key ^ 61) ^ (key >> 16)) & M
    key = (key + (key << 3)) & M
    key = key ^ (key >> 4)
    key = (key * 0x27D4EB2D) & M
    key = key ^ (key >> 15)
    return (key & 0x7FFFFFFF) % B


def revcomp(s):
    return "".join(COMP[c] for c in reversed(s))


def allowed(s):                                            # App. A.5 closed form of UTIL:46-75
    return "AA" not in s and not s.startswith("ACA")


def val(s):
    v = 0
    for c in s:
        v = v * 4 + "ACGT".index(c)
    return v


def norm(s):                                               # UTIL:77-100
    d = 4 ** len(s)
    r = revcomp(s)
    return min(val(s) if allowed(s) else d, val(r) if allowed(r) else d)


def records(fasta: str):
    recs, cur, seen = [], None, False
    for line in fasta.split("\n"):
        if line.startswith(">"):
            if cur is not None:
                recs.append(cur)
            cur, seen = "", True
        elif seen:
            cur += line
    if cur is not None:
        recs.append(cur)
    return recs


def count(fasta: str, k, m, max_b):
    """-> Counter {(bin, canonical kmer): count}"""
    B = min(4 ** m, max_b)
    out = Counter()
    for rec in records(fasta):
        for i in range(len(rec) - k + 1):
            w = rec[i:i + k]
            if any(c not in "ACGT" for c in w):
                continue
            sig = min(norm(w[j:j + m]) for j in range(k - m + 1))
            out[(hash_to_bucket(sig, B), min(w, revcomp(w)))] += 1
    return out
