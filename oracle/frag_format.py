"""frag_format.py — TEST INFRASTRUCTURE ONLY.  Reader / writer of the reference's fragment store
(<prefix>.sdx + <prefix>.frg, seq_db.rs:814-873 write_to_frag_files; frag_file_io.rs readers) in pure Python:
  * 7-byte tags "SDX:0.5" / "FRG:0.5";
  * bincode 2.0.0-rc (pgr-db/Cargo.toml) `config::standard()`: little endian, VARIABLE-length integers (u < 251 one byte;
    251 + u16, 252 + u32, 253 + u64), u8 and bool as one raw byte, Vec / String = varint length + items, enum = varint
    variant index, Option = 0 / 1 + value, usize as u64;
  * .sdx = (chunk_size: usize, Vec<(offset, length, bases)>, Vec<CompactSeq{source: Option<String>, name, id: u32,
    seq_frag_range: (u32, u32), len: usize}>);  .frg = raw-deflate stream per chunk of `chunk_size` fragments, each
    chunk = bincode Vec<Fragment>;
  * Fragment = AlnSegments((ref_frag_id: u32, reversed: bool, len: u32, Vec<AlnSegment>)) | Prefix(bytes) | Internal(bytes)
    | Suffix(bytes) (seq_db.rs:48-55); AlnSegment = FullMatch | Match(u32, u32) | Insertion(u8) (seq_db.rs:34-41).
Pinned by tests/test_frag_format.py: the reference's committed fixture test_seqs_frag.{sdx,frg} decodes to exactly the
66 sequences of test_seqs.fa.
"""
import zlib

FRAG_ALN, FRAG_PREFIX, FRAG_INTERNAL, FRAG_SUFFIX = 0, 1, 2, 3
SEG_FULL, SEG_MATCH, SEG_INS = 0, 1, 2


class Reader:
    def __init__(self, buf, pos=0):
        self.b, self.p = buf, pos

    def u8(self):
        v = self.b[self.p]
        self.p += 1
        return v

    def varint(self):
        t = self.u8()
        if t < 251:
            return t
        n = {251: 2, 252: 4, 253: 8, 254: 16}[t]
        v = int.from_bytes(self.b[self.p:self.p + n], "little")
        self.p += n
        return v

    def bytes_(self):
        n = self.varint()
        v = bytes(self.b[self.p:self.p + n])
        self.p += n
        return v

    def string(self):
        return self.bytes_().decode()


def enc_varint(v):
    if v < 251:
        return bytes([v])
    if v < 1 << 16:
        return b"\xfb" + v.to_bytes(2, "little")
    if v < 1 << 32:
        return b"\xfc" + v.to_bytes(4, "little")
    return b"\xfd" + v.to_bytes(8, "little")


def dec_fragment(r):
    tag = r.varint()
    if tag == FRAG_ALN:
        ref, rev, ln = r.varint(), bool(r.u8()), r.varint()
        segs = []
        for _ in range(r.varint()):
            st = r.varint()
            if st == SEG_FULL:
                segs.append((SEG_FULL,))
            elif st == SEG_MATCH:
                a = r.varint()
                segs.append((SEG_MATCH, a, r.varint()))
            else:
                segs.append((SEG_INS, r.u8()))
        return (FRAG_ALN, ref, rev, ln, segs)
    return (tag, r.bytes_())


def enc_fragment(f):
    if f[0] == FRAG_ALN:
        _, ref, rev, ln, segs = f
        out = [enc_varint(FRAG_ALN), enc_varint(ref), bytes([1 if rev else 0]), enc_varint(ln), enc_varint(len(segs))]
        for s in segs:
            out.append(enc_varint(s[0]))
            if s[0] == SEG_MATCH:
                out += [enc_varint(s[1]), enc_varint(s[2])]
            elif s[0] == SEG_INS:
                out.append(bytes([s[1]]))
        return b"".join(out)
    return enc_varint(f[0]) + enc_varint(len(f[1])) + bytes(f[1])


def enc_chunk(frags):
    return enc_varint(len(frags)) + b"".join(enc_fragment(f) for f in frags)


def read_sdx(path):
    b = open(path, "rb").read()
    assert b[:7] == b"SDX:0.5"
    r = Reader(b, 7)
    chunk_size = r.varint()
    addr = [(r.varint(), r.varint(), r.varint()) for _ in range(r.varint())]
    seqs = []
    for _ in range(r.varint()):
        source = r.string() if r.u8() else None
        name = r.string()
        sid = r.varint()
        rng = (r.varint(), r.varint())
        seqs.append({"source": source, "name": name, "id": sid, "seq_frag_range": rng, "len": r.varint()})
    assert r.p == len(b)
    return chunk_size, addr, seqs


def read_frg_chunks(path, addr):
    """-> the inflated bincode payload of every chunk"""
    b = open(path, "rb").read()
    assert b[:7] == b"FRG:0.5"
    return [zlib.decompress(b[7 + off:7 + off + ln], -15) for off, ln, _ in addr]


def decode_chunks(payloads):
    frags = []
    for p in payloads:
        r = Reader(p)
        for _ in range(r.varint()):
            frags.append(dec_fragment(r))
        assert r.p == len(p)
    return frags


COMP = {ord(a): ord(b) for a, b in zip("ACGTacgtNn", "TGCAtgcaNn")}


def reverse_complement(s):
    return bytes(COMP.get(c, c) for c in reversed(s))


def reconstruct_from_segs(base, segs):   # seq_db.rs:158-174
    out = bytearray()
    for s in segs:
        if s[0] == SEG_FULL:
            out += base
        elif s[0] == SEG_MATCH:
            out += base[s[1]:s[2]]
        else:
            out.append(s[1])
    return bytes(out)


def get_seq(frags, k, seq):   # seq_db.rs:685-735
    out = bytearray()
    a, n = seq["seq_frag_range"]
    for f in frags[a:a + n]:
        if f[0] in (FRAG_PREFIX, FRAG_SUFFIX):
            out += f[1]
        elif f[0] == FRAG_INTERNAL:
            out += f[1][k:]
        else:
            _, ref, rev, _ln, segs = f
            base = frags[ref]
            assert base[0] == FRAG_INTERNAL
            s = reconstruct_from_segs(base[1], segs)
            if rev:
                s = reverse_complement(s)
            out += s[k:]
    return bytes(out)
