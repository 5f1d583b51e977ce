"""`bio::io::fastq::{Reader, Record, Records}` (reference src/io/fastq.rs:140-330, 340-460): the sequential
reader.  The Writer is not part of the alignment path's I/O step and is not mirrored."""
from __future__ import annotations

from typing import Iterator, Optional

from .fasta import CheckError, _text_stream


class ReadError(OSError):
    """fastq.rs `ReadError`: MissingAt / IncompleteRecord / Io"""


class Record:
    """fastq.rs:340-460"""
    __slots__ = ("_id", "_desc", "_seq", "_qual")

    def __init__(self):
        self._id, self._desc, self._seq, self._qual = "", None, "", ""

    @staticmethod
    def new() -> "Record":
        return Record()

    @staticmethod
    def with_attrs(id: str, desc: Optional[str], seq: bytes, qual: bytes) -> "Record":
        r = Record()
        r._id, r._desc, r._seq, r._qual = id, desc, bytes(seq).decode("utf-8"), bytes(qual).decode("utf-8")
        return r

    def is_empty(self) -> bool:
        return not self._id and self._desc is None and not self._seq and not self._qual

    def check(self) -> None:
        """fastq.rs:388-410"""
        if not self._id:
            raise CheckError("EmptyId")
        if not self._seq.isascii():
            raise CheckError("NonAsciiSequence")
        if not all(("a" <= c <= "z") or ("A" <= c <= "Z") or c in "-.*" for c in self._seq):
            raise CheckError("InvalidSequence")
        if not self._qual.isascii():
            raise CheckError("NonAsciiQualities")
        if len(self.seq()) != len(self.qual()):
            raise CheckError("UnequalLength")

    def id(self) -> str:
        return self._id

    def desc(self) -> Optional[str]:
        return self._desc

    def seq(self) -> bytes:
        return self._seq.rstrip().encode("utf-8")

    def qual(self) -> bytes:
        return self._qual.rstrip().encode("utf-8")

    def clear(self) -> None:
        self._id, self._desc, self._seq, self._qual = "", None, "", ""

    def __str__(self) -> str:
        head = "@" + self._id + ((" " + self._desc) if self._desc is not None else "")
        return head + "\n" + self._seq + "\n+\n" + self._qual + "\n"


class Reader:
    """fastq.rs:151-305"""

    def __init__(self, stream):
        self._stream = stream

    @staticmethod
    def new(reader) -> "Reader":
        return Reader(_text_stream(reader))

    @staticmethod
    def from_file(path) -> "Reader":
        return Reader(open(path, "r", encoding="utf-8", newline=""))

    def read(self, record: Record) -> None:
        """fastq.rs:265-304 (multi-line sequences: as many quality lines as sequence lines are read)."""
        record.clear()
        line = self._stream.readline()
        if not line:
            return
        if not line.startswith("@"):
            raise ReadError("MissingAt")
        head = line[1:].rstrip()
        cut = head.find(" ")  # splitn(2, ' ')
        if cut < 0:
            record._id, record._desc = head, None
        else:
            record._id, record._desc = head[:cut], head[cut + 1:]
        line = self._stream.readline()
        seq, lines_read = [], 0
        while line and not line.startswith("+"):
            seq.append(line.rstrip())
            line = self._stream.readline()
            lines_read += 1
        record._seq = "".join(seq)
        qual = []
        for _ in range(lines_read):
            qual.append(self._stream.readline().rstrip())
        record._qual = "".join(qual)
        if not record._qual:
            raise ReadError("IncompleteRecord")

    def records(self) -> Iterator[Record]:
        while True:
            r = Record()
            self.read(r)
            if r.is_empty():
                return
            yield r
