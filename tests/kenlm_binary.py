"""TEST INFRASTRUCTURE ONLY: writes an ARPA model in the layout of a KenLM binary file (probing model type, format
version 5) -- an independent restatement, in Python, of the layout the C++ reader in csrc/b2c_lm_host.h parses
(kenlm's lm/binary_format.cc, lm/vocab.cc, lm/search_hashed.cc, util/probing_hash_table.hh as published; the kenlm
package itself is not available here, so files written by kenlm's own build_binary remain the unpinned case).

    write_probing_binary(arpa_path, out_path, multiplier=1.5, flip_signs=True)
"""
import struct

MASK = (1 << 64) - 1
MAGIC = b"mmap lm http://kheafield.com/code format version 5\n\0"


def murmur64a(data: bytes, seed: int = 0) -> int:
    m, r = 0xC6A4A7935BD1E995, 47
    n = len(data)
    h = (seed ^ (n * m)) & MASK
    full = n // 8 * 8
    for i in range(0, full, 8):
        k = struct.unpack_from("<Q", data, i)[0]
        k = (k * m) & MASK
        k ^= k >> r
        k = (k * m) & MASK
        h ^= k
        h = (h * m) & MASK
    tail = data[full:]
    if tail:
        for i in range(len(tail) - 1, -1, -1):
            h ^= tail[i] << (8 * i)
        h = (h * m) & MASK
    h ^= h >> r
    h = (h * m) & MASK
    h ^= h >> r
    return h


def combine(cur: int, nxt: int) -> int:
    return ((cur * 8978948897894561157) & MASK) ^ (((1 + nxt) * 17894857484156487943) & MASK)


def parse_arpa(path):
    grams, order, cur = {}, 0, 0
    with open(path, encoding="utf-8") as fh:
        for raw in fh:
            line = raw.strip()
            if not line:
                continue
            if line.startswith("\\"):
                if line.endswith("-grams:"):
                    cur = int(line[1:line.index("-")])
                    order = max(order, cur)
                    grams[cur] = []
                elif line.startswith("\\end\\"):
                    break
                continue
            if cur == 0:
                continue
            f = line.split()
            if len(f) < cur + 1:
                continue
            prob = float(f[0])
            words = f[1:1 + cur]
            backoff = float(f[1 + cur]) if len(f) > 1 + cur else None
            grams[cur].append((prob, words, backoff))
    return order, grams


def write_probing_binary(arpa_path, out_path, multiplier=1.5, flip_signs=True):
    order, grams = parse_arpa(arpa_path)
    assert order >= 2
    # word indices: <unk> is 0, the others 1.. in the order of the unigram section
    index = {"<unk>": 0}
    strings = ["<unk>"]
    for _, (w,), _ in grams[1]:
        if w not in index:
            index[w] = len(strings)
            strings.append(w)
    n_words = len(strings)
    counts = [len(grams[1]) + (0 if any(w[0] == "<unk>" for _, w, _ in grams[1]) else 1)] + [len(grams[n]) for n in range(2, order + 1)]
    assert counts[0] == n_words

    def key_of(words):
        ids = [index[w] for w in words]
        h = ids[-1]
        for c in reversed(ids[:-1]):
            h = combine(h, c)
        return h

    # contexts that a longer n-gram extends (for the no-extension mark of the backoff)
    extended = set()
    for n in range(2, order + 1):
        for _, words, _ in grams[n]:
            extended.add(tuple(words[:-1]))

    def backoff_bits(words, backoff, top):
        if top:
            return None
        b = 0.0 if backoff is None else backoff
        if b == 0.0:
            return struct.pack("<f", 0.0 if tuple(words) in extended else -0.0)
        return struct.pack("<f", b)

    def prob_bits(prob, i):
        p = abs(prob)
        # the sign bit of a stored prob is a flag of kenlm's left-state optimisation; readers take -|prob|
        return struct.pack("<f", p if (flip_signs and i % 3 == 0) else -p)

    def buckets(n):
        return max(n + 1, int(struct.unpack("<f", struct.pack("<f", multiplier))[0] * float(n)))

    out = bytearray()
    magic = MAGIC + b"\0" * ((-len(MAGIC)) % 8)
    assert len(magic) == 56
    out += magic
    out += struct.pack("<fffII", 0.0, 1.0, -0.5, 1, 0xFFFFFFFF)
    out += b"\0" * 4                                   # alignment of the u64
    out += struct.pack("<Q", 1)
    assert len(out) == 88
    out += struct.pack("<B3xfiB3xI", order, multiplier, 0, 1, 0)
    out += struct.pack("<%dQ" % order, *counts)
    out += b"\0" * ((-len(out)) % 8)
    # vocabulary
    vb = buckets(counts[0])
    table = [(0, 0)] * vb
    for w in range(1, n_words):
        k = murmur64a(strings[w].encode("utf-8"))
        b = k % vb
        while table[b][0] != 0:
            b = (b + 1) % vb
        table[b] = (k, w)
    out += struct.pack("<II", 0, n_words)
    for k, v in table:
        out += struct.pack("<QI4x", k, v)
    # unigrams by index, one extra slot
    uni = [struct.pack("<ff", -100.0, -0.0)] * (counts[0] + 1)
    for i, (prob, (w,), backoff) in enumerate(grams[1]):
        uni[index[w]] = prob_bits(prob, i) + backoff_bits([w], backoff, False)
    out += b"".join(uni)
    # middle orders and the longest order
    for n in range(2, order + 1):
        top = n == order
        nb = buckets(counts[n - 1])
        width = 12 if top else 16
        tab = [None] * nb
        for i, (prob, words, backoff) in enumerate(grams[n]):
            k = key_of(words)
            assert k != 0
            b = k % nb
            while tab[b] is not None:
                assert tab[b][0] != k, "duplicate n-gram"
                b = (b + 1) % nb
            tab[b] = (k, (struct.pack("<f", -abs(prob)) if top else prob_bits(prob, i)) + (b"" if top else backoff_bits(words, backoff, False)))
        for e in tab:
            out += (b"\0" * width) if e is None else struct.pack("<Q", e[0]) + e[1]
    for s in strings:
        out += s.encode("utf-8") + b"\0"
    with open(out_path, "wb") as fh:
        fh.write(bytes(out))
    return {"order": order, "counts": counts, "words": n_words}
