"""Adversarial random tiles built directly as md_reads_soa (no BAM involved): random CIGARs with every operator
(M I D N S H P = X), zero-length and inconsistent records, odd flag combinations, duplicate and colliding names,
reads hanging over the contig and tile edges, all phred encodings.  Used by the CPU tier (structure only) and the
GPU tier (CUDA path == oracle port, record for record)."""
import ctypes as C

import numpy as np

from methyldackel_b200 import _abi as A


class Tile:
    """Keeps the numpy arrays alive behind an MdReadsSoa."""

    def __init__(self, rng, reflen, n, qual_bits=8, maxlen=180, depth_hot=False, dup_names=0.02):
        self.keep = []
        pos = np.sort(rng.integers(-5 if False else 0, max(1, reflen - 20), size=n)).astype(np.int32)
        if depth_hot:
            pos = np.sort(np.clip(rng.normal(reflen / 2, 300, size=n), 0, reflen - 20)).astype(np.int32)
        flags, mapq, aux, lq, cig_off, cig, seqw, qualw, seq_off, qual_off, keys = [], [], [], [], [0], [], [], [], [], [], []
        alphabet = {2: [2, 12, 23, 37], 4: list(range(5, 21)), 8: list(range(0, 60)) + [255]}[qual_bits]
        lut = (alphabet + [0] * 16)[:16] if qual_bits != 8 else [0] * 16
        for i in range(n):
            style = rng.random()
            L = int(rng.integers(1, maxlen))
            ops = []
            if style < 0.55:
                ops = [(0, L)]
            elif style < 0.9:
                # random legal-looking CIGAR whose query length is L
                left = L
                if rng.random() < 0.3:
                    ops.append((5, int(rng.integers(1, 9))))
                if rng.random() < 0.4 and left > 2:
                    k = int(rng.integers(1, min(20, left))); ops.append((4, k)); left -= k
                while left > 0:
                    op = int(rng.choice([0, 0, 0, 7, 8, 1, 2, 3, 6]))
                    k = int(rng.integers(1, 40))
                    if op in (0, 7, 8, 1):
                        k = min(k, left); left -= k
                    ops.append((op, k))
                if rng.random() < 0.3:
                    ops.append((4, 0)); ops.append((5, 3))
            elif style < 0.95:
                ops = [(4, L)]                      # no reference span at all
            else:
                ops = [(0, L + int(rng.integers(1, 5)))]   # CIGAR does not match l_qseq
            f = int(rng.choice([99, 147, 83, 163, 0, 16, 1, 65, 129, 113, 177, 97, 145, 67, 131, 73, 89, 137, 153, 4, 1024 + 99, 512 + 147, 256 + 83, 2048 + 163, 77, 141]))
            if rng.random() < 0.1:
                f = int(rng.integers(0, 4096))
            flags.append(f); mapq.append(int(rng.choice([0, 3, 9, 10, 11, 30, 60, 255]))); aux.append(int(rng.choice([0, 0, 0, 1, 2, 3, 4, 5, 6])))
            lq.append(L)
            for op, k in ops:
                cig.append((k << 4) | op)
            cig_off.append(len(cig))
            bases = rng.choice([1, 2, 4, 8, 15, 3, 0], size=L, p=[0.22, 0.25, 0.25, 0.22, 0.03, 0.02, 0.01]).astype(np.uint8)
            packed = np.zeros(((L + 1) // 2 + 3) // 4 * 4, dtype=np.uint8)
            packed[: (L + 1) // 2] = (np.pad(bases, (0, L % 2))[0::2] << 4) | np.pad(bases, (0, L % 2))[1::2]
            seq_off.append(sum(len(x) for x in seqw) // 4); seqw.append(packed)
            q = rng.choice(alphabet, size=L).astype(np.uint8)
            if qual_bits == 8:
                qp = np.zeros((L + 7) // 8 * 8, dtype=np.uint8); qp[:L] = q
            else:
                codes = np.array([alphabet.index(int(v)) for v in q], dtype=np.uint64)
                words = np.zeros((L * qual_bits + 63) // 64, dtype=np.uint64)
                per = 64 // qual_bits
                for j, cde in enumerate(codes):
                    words[j // per] |= np.uint64(int(cde) << ((j % per) * qual_bits))
                qp = words.view(np.uint8)
            qual_off.append(sum(len(x) for x in qualw) // 8); qualw.append(qp)
            keys.append(int(rng.integers(1, 2 ** 63)))
        keys = np.array(keys, dtype=np.uint64)
        # mates: give pairs of records the same name; a few names three or four times; one key == 0
        idx = rng.permutation(n)
        for a, b in zip(idx[0: n // 2: 2], idx[1: n // 2: 2]):
            keys[b] = keys[a]
        for t in range(int(n * dup_names)):
            keys[int(rng.integers(0, n))] = keys[int(rng.integers(0, n))]
        if n:
            keys[int(rng.integers(0, n))] = 0
        arr = lambda x, dt: np.ascontiguousarray(np.array(x, dtype=dt))
        self.pos = pos; self.flag = arr(flags, np.uint16); self.mapq = arr(mapq, np.uint8); self.aux = arr(aux, np.uint8); self.lq = arr(lq, np.uint32)
        self.cig_off = arr(cig_off, np.uint32); self.seq_off = arr(seq_off, np.uint32); self.qual_off = arr(qual_off, np.uint32); self.keys = keys
        self.cig = arr(cig if cig else [0], np.uint32)
        self.seq = np.ascontiguousarray(np.concatenate(seqw + [np.zeros(64, np.uint8)])); self.qual = np.ascontiguousarray(np.concatenate(qualw + [np.zeros(64, np.uint8)]))
        s = A.MdReadsSoa()
        s.n_reads = n; s.n_cigar_ops = len(cig); s.seq_words = sum(len(x) for x in seqw) // 4; s.qual_words = sum(len(x) for x in qualw) // 8
        p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        s.pos = p(self.pos, C.c_int32); s.flag = p(self.flag, C.c_uint16); s.mapq = p(self.mapq, C.c_uint8); s.aux = p(self.aux, C.c_uint8); s.l_qseq = p(self.lq, C.c_uint32)
        s.cigar_off = p(self.cig_off, C.c_uint32); s.seq_off = p(self.seq_off, C.c_uint32); s.qual_off = p(self.qual_off, C.c_uint32); s.frag_key = p(self.keys, C.c_uint64)
        s.cigar = p(self.cig, C.c_uint32); s.seq = p(self.seq, C.c_uint32); s.qual = p(self.qual, C.c_uint64)
        s.qual_bits = qual_bits
        for k in range(16):
            s.qual_lut[k] = lut[k] & 0xff
        self.soa = s


def random_reference(rng, n):
    r = rng.choice(list(b"ACGTacgtNn"), size=n, p=[0.22, 0.24, 0.24, 0.22, 0.01, 0.02, 0.02, 0.01, 0.015, 0.005]).astype(np.uint8)
    return bytes(r)


def random_config(rng):
    kw = dict(keepCpG=int(rng.random() < 0.8), keepCHG=int(rng.random() < 0.6), keepCHH=int(rng.random() < 0.6), minMapq=int(rng.choice([0, 5, 10, 40])),
              minPhred=int(rng.choice([1, 5, 13, 30])), keepDupes=int(rng.random() < 0.3), keepSingleton=int(rng.random() < 0.3), keepDiscordant=int(rng.random() < 0.5),
              ignoreFlags=int(rng.choice([0xF00, 0, 0x900, 0x400])), requireFlags=int(rng.choice([0, 0, 2, 0x40])), ignoreNH=int(rng.random() < 0.3))
    if not (kw["keepCpG"] or kw["keepCHG"] or kw["keepCHH"]):
        kw["keepCpG"] = 1
    if rng.random() < 0.4:
        kw["minOppositeDepth"] = int(rng.integers(1, 4)); kw["maxVariantFrac"] = float(rng.choice([0.0, 0.1, 0.5, 1.0]))
    if rng.random() < 0.4:
        kw["bounds"] = [int(x) for x in rng.integers(0, 40, size=16)]
    if rng.random() < 0.4:
        kw["absoluteBounds"] = [int(x) for x in rng.integers(0, 25, size=16)]
    if rng.random() < 0.15:
        kw["minConversionEfficiency"] = float(rng.choice([0.5, 0.9, 1.0]))
    return A.default_config(**kw)
