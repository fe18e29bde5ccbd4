"""Pure-Python restatement of the reference's FASTA / FASTQ record readers (pgr-db/src/fasta_io.rs:46-172).
Test infrastructure only: the product parser is pgr_tk_b200/host/fastx_ingest.cpp."""


def _read_until(buf, pos, byte):
    """BufRead::read_until: bytes up to and including `byte` (or to the end) -> (chunk, new position)"""
    i = buf.find(bytes([byte]), pos)
    end = len(buf) if i < 0 else i + 1
    return buf[pos:end], end


def _take_id(line):
    head, _ = _read_until(line, 0, ord(" "))
    return bytes(c for c in head if c not in (0x0A, 0x20, 0x0D))


def parse_fasta(buf):
    """fasta_io.rs:46-118: the constructor consumes one byte; a record = header line, then every byte up to the next '>'
    minus '\\n' '\\r' '>'; `None` when the header read returns 0 bytes"""
    if len(buf) == 0:
        raise IOError("empty file")                   # fasta_io.rs:58-63
    pos, out = 1, []
    while True:
        line, pos2 = _read_until(buf, pos, 0x0A)
        if len(line) == 0:
            return out
        pos = pos2
        chunk, pos = _read_until(buf, pos, ord(">"))
        out.append((_take_id(line), bytes(c for c in chunk if c not in (0x0A, 0x0D, ord(">")))))


def parse_fastq(buf):
    """fasta_io.rs:120-165: id line, ONE sequence line, read_until('+'), two read_until('\\n'), read_until('@'); when the last
    call returns 0 bytes the function returns None WITHOUT yielding the record it has just read"""
    if len(buf) == 0:
        raise IOError("empty file")
    pos, out = 1, []
    while True:
        line, pos = _read_until(buf, pos, 0x0A)
        rid = _take_id(line)
        seq, pos = _read_until(buf, pos, 0x0A)
        seq = bytes(c for c in seq if c not in (0x0A, 0x0D))
        _, pos = _read_until(buf, pos, ord("+"))
        _, pos = _read_until(buf, pos, 0x0A)
        _, pos = _read_until(buf, pos, 0x0A)
        last, pos = _read_until(buf, pos, ord("@"))
        if len(last) == 0:
            return out
        out.append((rid, seq))


def parse_fastx(buf):
    return parse_fastq(buf) if buf[:1] == b"@" else parse_fasta(buf)
