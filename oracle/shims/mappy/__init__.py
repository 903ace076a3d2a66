"""Import shim (test infrastructure): minimal stand-in for mappy (I/O only, no arithmetic).

The reference hot path only calls `mappy.fastx_read` (boss/runs/reference.py:328); `Aligner`
is constructed by boss/mapper.py but never used for mapping by the golden-vector generator.
"""
import gzip
from pathlib import Path


def _open(path):
    path = str(path)
    return gzip.open(path, "rt") if path.endswith(".gz") else open(path, "r")


def fastx_read(path, read_comment=False):
    with _open(path) as fh:
        first = fh.read(1)
        if not first:
            return
        fh.seek(0)
        if first == ">":
            name, comment, chunks = None, None, []
            for line in fh:
                if line.startswith(">"):
                    if name is not None:
                        yield (name, "".join(chunks), None, comment) if read_comment else (name, "".join(chunks), None)
                    head = line[1:].rstrip("\n").split(None, 1)
                    name = head[0] if head else ""
                    comment = head[1] if len(head) > 1 else None
                    chunks = []
                else:
                    chunks.append(line.strip())
            if name is not None:
                yield (name, "".join(chunks), None, comment) if read_comment else (name, "".join(chunks), None)
        else:
            while True:
                head = fh.readline()
                if not head:
                    break
                seq = fh.readline().rstrip("\n")
                fh.readline()
                qual = fh.readline().rstrip("\n")
                parts = head[1:].rstrip("\n").split(None, 1)
                name = parts[0]
                comment = parts[1] if len(parts) > 1 else None
                yield (name, seq, qual, comment) if read_comment else (name, seq, qual)


class Aligner:
    def __init__(self, seq=None, fn_idx_in=None, fn_idx_out=None, **kwargs):
        if fn_idx_out:
            Path(fn_idx_out).touch()

    def __bool__(self):
        return True

    def map(self, *args, **kwargs):
        return iter(())


class ThreadBuffer:
    pass
