"""literal.py — line-by-line transliteration of the reference's Scala for the k-mer counting path.

TEST INFRASTRUCTURE ONLY (never imported by fastkmer_b200/).  Where oracle/fkm_oracle.cpp restates the
algorithm with its own data layout, this file keeps the reference's: `Kmer` with 31 nucleotides per 64-bit
Long (UTIL:17,138-172), all code paths of `Kmer.readFromKmer` (UTIL:174-295), `Mmer` (UTIL:511-560), `RIndex`
and `priorityQueueWithIndexes` (UTIL:562-681), `getSuperKmers` (SBKC:34-169), `extractKXmers` (SBKC:428-660) and
`extractKXmersHT` (SBKC:664-739).  JVM arithmetic is emulated: Long/Int wrap, `>>` is arithmetic, shift counts
are taken modulo 64 (Long) / 32 (Int).  Scala precedence: `*` binds tighter than `+ -`, which bind tighter than
`<< >>`, which bind tighter than `&`, `^`, `|` — parenthesised here accordingly.

UTIL = src/main/scala/skc/package.scala, SBKC = src/main/scala/skc/SparkBinKmerCounter.scala.
Slow (pure Python): use on inputs of a few thousand k-mers.
"""
import heapq
import math

M64 = (1 << 64) - 1
M32 = (1 << 32) - 1


def s64(x):
    x &= M64
    return x - (1 << 64) if x >> 63 else x


def s32(x):
    x &= M32
    return x - (1 << 32) if x >> 31 else x


def shl64(x, n):
    return s64((x & M64) << (n & 63))


def shr64(x, n):                      # arithmetic shift of a JVM Long
    return s64(x) >> (n & 63)


# ------------------------------------------------------------------ UTIL:17-44
nucleotidesPerLong = 31
nucleotideBitmasks = {ord("A"): 0, ord("C"): 1, ord("G"): 2, ord("T"): 3}
nucleotideReprByte = b"ACGT"
nucleotideRC = [3, 2, 1, 0]
upper2bitMask = int(math.pow(2, 2 * nucleotidesPerLong)) - 1


def is_allowed(_mmer, length):        # UTIL:46-75
    mmer = _mmer
    for _ in range(0, length - 3):
        if (mmer & 0xF) == 0:
            return False
        mmer >>= 2
    if mmer == 0:
        return False
    if mmer == 0x04:
        return False
    if (mmer & 0x3C) == 0:
        return False
    if (mmer & 0xF) == 0:
        return False
    return True


def reverse_complement(seq, length):  # UTIL:103-115
    curSeq, rev, shift = seq, 0, length * 2 - 2
    for _ in range(length):
        rev = s64(rev + shl64(3 - (curSeq & 3), shift))
        curSeq = shr64(curSeq, 2)
        shift -= 2
    return rev


def fillNorm(sigLen):                 # UTIL:77-100
    default_signature = s32(1 << ((sigLen * 2) & 31))
    norm = [0] * (1 << (sigLen * 2))
    for i in range(default_signature):
        rev = s32(reverse_complement(i, sigLen))
        str_val = i if is_allowed(i, sigLen) else default_signature
        rev_val = rev if is_allowed(rev, sigLen) else default_signature
        norm[i] = min(str_val, rev_val)
    return norm


class LazyNorm:
    """norm(i) of UTIL:77-100 computed on demand: the table has 4^m entries (67 M for m=13), far too many to
    fill in pure Python.  Entry i is exactly what fillNorm would have stored."""

    def __init__(self, sigLen):
        self.sigLen, self.default_signature, self.cache = sigLen, s32(1 << ((sigLen * 2) & 31)), {}

    def __getitem__(self, i):
        v = self.cache.get(i)
        if v is None:
            rev = s32(reverse_complement(i, self.sigLen))
            str_val = i if is_allowed(i, self.sigLen) else self.default_signature
            rev_val = rev if is_allowed(rev, self.sigLen) else self.default_signature
            v = self.cache[i] = min(str_val, rev_val)
        return v


def fillRightOnesMask(runlength):     # UTIL:118-124
    res = 0
    for i in range(runlength):
        res |= shl64(1, i)
    return s64(res)


class Kmer:                           # UTIL:138-503
    def __init__(self, length):
        self.length = length
        self._data = None

    @staticmethod
    def fromBytes(length, s, offset):                         # UTIL:303-307
        k = Kmer(length)
        k._data = k._readKmer(length, s, offset)
        return k

    @staticmethod
    def fromKmer(length, frm, startPos, endPos, orientation):  # UTIL:299-302
        k = Kmer(length)
        k._data = k.readFromKmer(frm, startPos, endPos, length, orientation)
        return k

    def _readKmer(self, length, s, offset):                   # UTIL:144-172
        data = [0] * int(math.ceil(length / nucleotidesPerLong))
        currentLng, slice_, i = 0, 0, 0
        while i < length:
            currentLng = shl64(currentLng, 2)
            currentLng |= nucleotideBitmasks[s[offset + i]]
            i += 1
            if i % nucleotidesPerLong == 0 or i == length:
                currentLng &= upper2bitMask
                data[slice_] = currentLng
                currentLng = 0
                slice_ += 1
        return data

    def readFromKmer(self, fromKmer, startPos, endPos, amt, orientation):   # UTIL:174-295
        length = self.length
        data = [0] * int(math.ceil(amt / nucleotidesPerLong))
        fromEndSlice, fromEndOffset = fromKmer.getSliceOffset(endPos)
        assert (fromEndSlice, fromEndOffset) != (-1, -1), "endPos is invalid"
        fromStartSlice, fromStartOffset = fromKmer.getSliceOffset(startPos)
        assert (fromStartSlice, fromStartOffset) != (-1, -1), "startPos is invalid"
        excess = fromKmer.length % nucleotidesPerLong
        toExcess = amt % nucleotidesPerLong
        origFinalPadding = (nucleotidesPerLong - excess) * 2 if excess > 0 else 0
        if orientation == 0:
            slice_ = 0
            if fromStartOffset == 0:                                           # UTIL:198-209
                while fromEndSlice - fromStartSlice >= 0:
                    data[slice_] = fromKmer._data[fromStartSlice]
                    fromStartSlice += 1
                    slice_ += 1
                data[len(data) - 1] = shr64(data[len(data) - 1], 2 * (nucleotidesPerLong - 1) - fromEndOffset)
            elif fromStartSlice == fromEndSlice:                               # UTIL:210-214
                data[0] = shr64(fromKmer._data[fromStartSlice], nucleotidesPerLong * 2 - fromEndOffset - 2) & fillRightOnesMask(length * 2)
            else:                                                              # UTIL:215-253
                curOffset = fromStartOffset
                while slice_ < len(data) - 1:
                    data[slice_] = shl64(fromKmer._data[fromStartSlice], curOffset) & upper2bitMask
                    fromStartSlice += 1
                    curLng = fromKmer._data[fromStartSlice]
                    if fromStartSlice == len(fromKmer._data) - 1:
                        curLng = shr64(curLng, nucleotidesPerLong * 2 - curOffset - origFinalPadding)
                    else:
                        curLng = shr64(curLng, nucleotidesPerLong * 2 - curOffset)
                    data[slice_] |= curLng
                    slice_ += 1
                if fromStartSlice != fromEndSlice:
                    pad = origFinalPadding if fromEndSlice == len(fromKmer._data) - 1 else 0
                    data[slice_] = shl64(fromKmer._data[fromStartSlice] & fillRightOnesMask(nucleotidesPerLong * 2 - curOffset),
                                         fromEndOffset + 2 - pad)
                data[slice_] |= shr64(fromKmer._data[fromEndSlice], (nucleotidesPerLong - 1) * 2 - fromEndOffset)
                if toExcess != 0:
                    data[slice_] &= fillRightOnesMask(toExcess * 2)
        else:                                                                  # UTIL:256-292
            slice_ = 0
            curLng = shr64(fromKmer._data[fromEndSlice], (nucleotidesPerLong - 1) * 2 - fromEndOffset)
            written, written_this_slice = 0, 0
            if fromEndSlice != len(fromKmer._data) - 1:
                lastSliceQuantity = fromEndOffset // 2 + 1
            else:
                lastSliceQuantity = fromEndOffset // 2 - origFinalPadding // 2 + 1
            while written < length:
                data[slice_] = shl64(data[slice_], 2)
                data[slice_] = s64(data[slice_] + nucleotideRC[curLng & 3])
                written += 1
                written_this_slice += 1
                curLng = shr64(curLng, 2)
                if written == lastSliceQuantity or written_this_slice == nucleotidesPerLong:
                    fromEndSlice -= 1
                    if fromEndSlice >= 0:
                        curLng = fromKmer._data[fromEndSlice]
                    written_this_slice = 0
                if written % nucleotidesPerLong == 0:
                    slice_ += 1
        return data

    def lastM(self, mMask, norm, m=0):                         # UTIL:310-326
        if len(self._data) == 1 or self.length % nucleotidesPerLong >= m:
            return norm[s32(self._data[len(self._data) - 1] & mMask)]
        res = 0
        for i in range(self.length - m, self.length):
            res = s32(res << 2)
            res |= self.getNumSymbol(i)
        return norm[res]

    def firstM(self, m):                                       # UTIL:329-334
        if len(self._data) > 1 or self.length == nucleotidesPerLong:
            return shr64(self._data[0], (nucleotidesPerLong - m) * 2)
        return shr64(self._data[0], ((self.length % nucleotidesPerLong) - m) * 2)

    def getSignature(self, sigLen, norm):                      # UTIL:337-357
        mmer = Mmer(sigLen, 0, norm)
        pos = 0
        for i in range(sigLen):
            mmer.insert(self.getNumSymbol(i))
        sig = mmer.get()
        for i in range(sigLen, self.length):
            mmer.insert(self.getNumSymbol(i))
            if mmer.get() < sig:
                sig = mmer.get()
                pos = i - sigLen + 1
        return sig, pos

    def getSliceOffset(self, pos):                             # UTIL:360-373
        if pos >= self.length:
            return -1, -1
        slice_ = pos // nucleotidesPerLong
        if slice_ == len(self._data) - 1:
            offset = (nucleotidesPerLong - (self.length - pos)) * 2
        else:
            offset = (pos % nucleotidesPerLong) * 2
        return slice_, offset

    def getNumSymbol(self, pos):                               # UTIL:376-386
        slice_, offset = self.getSliceOffset(pos)
        if slice_ == -1:
            return -1
        mask = shr64(upper2bitMask, offset)
        symbol = self._data[slice_] & mask
        symbol = shr64(symbol, (nucleotidesPerLong * 2) - offset - 2)
        return symbol

    def compare(self, that):                                   # UTIL:389-404 (slice-wise signed Long compare)
        assert self.length == that.length
        for a, b in zip(self._data, that._data):
            if a != b:
                return -1 if a < b else 1
        return 0

    def key(self):                                             # equals/hashCode identity, UTIL:406-413
        return tuple(self._data)

    def toByteArray(self):                                     # UTIL:416-454
        result = bytearray(self.length)
        slice_ = len(self._data) - 1
        excess = self.length % nucleotidesPerLong
        i, j = 0, self.length - 1
        amt = excess if excess > 0 else nucleotidesPerLong
        currentLng = self._data[slice_]
        while j >= 0:
            binNucleotide = currentLng & 3
            currentLng = shr64(currentLng, 2)
            result[j] = nucleotideReprByte[binNucleotide]
            j -= 1
            i += 1
            if i == amt:
                slice_ -= 1
                i = 0
                amt = nucleotidesPerLong
                if slice_ >= 0:
                    currentLng = self._data[slice_]
        return bytes(result)

    def __str__(self):                                         # UTIL:496-500
        return self.toByteArray().decode()


class Mmer:                                                    # UTIL:511-560
    def __init__(self, length, seq, norm):
        self.mask = s32((1 << ((length * 2) & 31)) - 1)
        self._data = seq
        self.norm = norm
        self.currentVal = norm[seq]

    def get(self):
        return self.currentVal

    def insert(self, symb):
        self._data = s32(self._data << 2)
        self._data = s32(self._data + symb)
        self._data &= self.mask
        self.currentVal = self.norm[self._data]


class RIndex:                                                  # UTIL:562-601
    def __init__(self, arr, startPos, endPos, shift, kmer_length):
        self.arr, self.endPos, self.shift, self.kmer_length = arr, endPos, shift, kmer_length
        self.currentPos = startPos
        self.pointedKmer = None
        self._readKmer()

    def advance(self):
        self.currentPos += 1
        if not self.exhausted():
            self._readKmer()

    def _readKmer(self):
        nextKmer = self.arr[self.currentPos]
        self.pointedKmer = Kmer.fromKmer(self.kmer_length, nextKmer, self.shift, self.shift + self.kmer_length - 1, 0)

    def exhausted(self):
        return self.currentPos >= self.endPos


def priorityQueueWithIndexes(arr, k):                          # UTIL:642-681 (PointedMinOrder: a min-heap on the pointed k-mer)
    heap, tick = [], [0]

    def enqueue(r):
        tick[0] += 1
        heapq.heappush(heap, (r.pointedKmer.key(), tick[0], r))

    for r_index in range(len(arr)):
        a = arr[r_index]
        if len(a) > 0:
            starts = [0] * r_index
            enqueue(RIndex(a, 0, len(a), 0, k))
            if len(a) > 1:
                for j in range(1, len(a)):
                    for i in range(0, r_index):
                        if a[j - 1].firstM(i + 1) != a[j].firstM(i + 1):
                            enqueue(RIndex(a, starts[i], j, i + 1, k))
                            starts[i] = j
            for i in range(len(starts)):
                enqueue(RIndex(a, starts[i], len(a), i + 1, k))
    return heap, enqueue


def hash_to_bucket(s, B):                                      # UTIL:686-695
    key = s32(s)
    c2 = 0x27D4EB2D
    key = s32((key ^ 61) ^ ((key & M32) >> 16))
    key = s32(key + s32(key << 3))
    key = s32(key ^ ((key & M32) >> 4))
    key = s32(key * c2)
    key = s32(key ^ ((key & M32) >> 15))
    return (key & 0x7FFFFFFF) % B


def notANucleotide(c):                                         # UTIL:697
    return not (c == ord("A") or c == ord("C") or c == ord("G") or c == ord("T"))


def getOrientation(s, i, j):                                   # UTIL:721-728
    while True:
        start, end = s.getNumSymbol(i), s.getNumSymbol(j)
        if start < nucleotideRC[end]:
            return 0
        if start > nucleotideRC[end] or i >= j:
            return 1
        i, j = i + 1, j - 1


def firstAndLastOccurrenceOfInvalidNucleotide(s, start, end):  # UTIL:739-754
    first, last = -1, -1
    for i in range(start, end):
        if notANucleotide(s[i]):
            if first == -1:
                first = i - start
                last = i - start
            else:
                last = i - start
    return first, last


# ------------------------------------------------------------------ SBKC:34-169
def getSuperKmers(k, m, B, reads):
    """reads: iterable of record values (bytes, newlines already inside as in FASTdoop's getValue)."""
    def bin_(s):
        return hash_to_bucket(s, B)
    out = [[] for _ in range(B)]
    norm = fillNorm(m) if m <= 7 else LazyNorm(m)                                # SBKC:47 (same entries, computed on demand)
    lastMmask = s64((1 << ((m * 2) & 31)) - 1)
    for read in reads:
        cur = read.replace(b"\n", b"")                                           # SBKC:63-64
        if len(cur) >= k:
            min_value, min_pos = -1, -1
            super_kmer_start = 0
            i = 0
            while i < len(cur) - k + 1:
                N_pos = firstAndLastOccurrenceOfInvalidNucleotide(cur, i, i + k)
                if N_pos[0] != -1:
                    if super_kmer_start < i:
                        out[bin_(min_value)].append(Kmer.fromBytes(i - 1 + k - super_kmer_start, cur, super_kmer_start))
                    super_kmer_start = i + N_pos[1] + 1
                    i += N_pos[1] + 1
                else:
                    s = Kmer.fromBytes(k, cur, i)
                    if i > min_pos:
                        if super_kmer_start < i:
                            out[bin_(min_value)].append(Kmer.fromBytes(i - 1 + k - super_kmer_start, cur, super_kmer_start))
                            super_kmer_start = i
                        sig, pos = s.getSignature(m, norm)
                        min_value, min_pos = sig, pos + i
                    else:
                        last = s.lastM(lastMmask, norm, m)
                        if last < min_value:
                            if super_kmer_start < i:
                                out[bin_(min_value)].append(Kmer.fromBytes(i - 1 + k - super_kmer_start, cur, super_kmer_start))
                                super_kmer_start = i
                            min_value, min_pos = last, i + k - m
                    i += 1
            if len(cur) - super_kmer_start >= k:
                N_pos = firstAndLastOccurrenceOfInvalidNucleotide(cur, i, len(cur))
                if N_pos[0] == -1:
                    out[bin_(min_value)].append(Kmer.fromBytes(len(cur) - super_kmer_start, cur, super_kmer_start))
                elif i + N_pos[0] >= super_kmer_start + k:
                    out[bin_(min_value)].append(Kmer.fromBytes(i + N_pos[0], cur, super_kmer_start))
    return {b: arr for b, arr in enumerate(out) if arr}


# ------------------------------------------------------------------ SBKC:428-660
def extractKXmers(k, x, bins):
    """bins: {bin: [super-k-mer Kmer]} -> {bin: [(kmer string, count)] in file order}"""
    result = {}
    for bin_no, superKmers in bins.items():
        unsortedR = [[] for _ in range(x + 1)]
        for sk in superKmers:
            lastOrientation, orientation, runLength, runStart = -1, -1, 0, 0
            for i in range(0, sk.length - k + 1):
                orientation = getOrientation(sk, i, i + k - 1)
                if orientation == lastOrientation:
                    runLength += 1
                    if runLength == x + 1:
                        unsortedR[runLength - 1].append(Kmer.fromKmer(k + runLength - 1, sk, runStart, runStart + k + runLength - 2, orientation))
                        runLength = 0
                        runStart = i
                        lastOrientation = -1
                else:
                    if lastOrientation != -1:
                        unsortedR[runLength - 1].append(Kmer.fromKmer(k + runLength - 1, sk, runStart, runStart + k + runLength - 2, lastOrientation))
                    runLength = 1
                    runStart = i
                    lastOrientation = orientation
            if runLength > 0:
                unsortedR[runLength - 1].append(Kmer.fromKmer(k + runLength - 1, sk, runStart, runStart + k + runLength - 2, lastOrientation))
        sortedR = [sorted(a, key=Kmer.key) for a in unsortedR]                   # SBKC:540-542 (Kmer.compare order)
        heap, enqueue = priorityQueueWithIndexes(sortedR, k)
        lines = []
        last_kmer, last_kmer_cnt = None, 0
        while heap:
            _, _, index = heapq.heappop(heap)
            if last_kmer is not None and index.pointedKmer.key() == last_kmer.key():
                last_kmer_cnt += 1
            else:
                if last_kmer is not None:
                    lines.append((str(last_kmer), last_kmer_cnt))
                last_kmer = index.pointedKmer
                last_kmer_cnt = 1
            index.advance()
            if not index.exhausted():
                enqueue(index)
        if last_kmer is not None:
            lines.append((str(last_kmer), last_kmer_cnt))
        if lines:
            result[bin_no] = lines
    return result


# ------------------------------------------------------------------ SBKC:664-739
def extractKXmersHT(k, bins):
    result = {}
    for bin_no, superKmers in bins.items():
        table = {}
        for sk in superKmers:
            for i in range(0, sk.length - k + 1):
                orientation = getOrientation(sk, i, i + k - 1)
                kmer = Kmer.fromKmer(k, sk, i, i + k - 1, orientation)
                key = kmer.key()
                if key in table:
                    table[key][1] += 1
                else:
                    table[key] = [kmer, 1]
        if table:
            result[bin_no] = [(str(km), c) for km, c in table.values()]
    return result


def records(fasta: bytes):
    """Record values as FASTdoop would hand them over (sequence lines, newlines kept): SURVEY App. A.1."""
    recs, cur, seen = [], None, False
    for line in fasta.split(b"\n"):
        if line.startswith(b">"):
            if cur is not None:
                recs.append(cur)
            cur, seen = b"", True
        elif seen:
            cur += line + b"\n"
    if cur is not None:
        recs.append(cur)
    return recs


def count(fasta: bytes, k, m, x, max_b, use_ht):
    """-> sorted list of (bin, kmer string, count)"""
    B = int(min(math.pow(4, m), max_b))                                          # TCFG:32
    bins = getSuperKmers(k, m, B, records(fasta))
    per_bin = extractKXmersHT(k, bins) if use_ht else extractKXmers(k, x, bins)
    return sorted((b, s, c) for b, lines in per_bin.items() for s, c in lines), per_bin
