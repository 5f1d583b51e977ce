"""`bio::io::fasta::{Reader, Record, Records}` (reference src/io/fasta.rs:168-360, 943-1170): the sequential
reader.  Index / IndexedReader / Writer are not part of the alignment path's I/O step and are not mirrored."""
from __future__ import annotations

import io
from typing import Iterator, Optional


class CheckError(ValueError):
    """fasta.rs `CheckError`: EmptyId / NonAsciiSequence / InvalidSequence"""


class Record:
    """fasta.rs:943-1035"""
    __slots__ = ("_id", "_desc", "_seq")

    def __init__(self):
        self._id, self._desc, self._seq = "", None, ""

    @staticmethod
    def new() -> "Record":
        return Record()

    @staticmethod
    def with_attrs(id: str, desc: Optional[str], seq: bytes) -> "Record":
        r = Record()
        r._id, r._desc, r._seq = id, desc, bytes(seq).decode("utf-8")
        return r

    def is_empty(self) -> bool:
        return not self._id and self._desc is None and not self._seq

    def check(self) -> None:
        """fasta.rs:993-1009: raises CheckError where the reference returns Err."""
        if not self._id:
            raise CheckError("EmptyId")
        if not self._seq.isascii():
            raise CheckError("NonAsciiSequence")
        if not all(("a" <= c <= "z") or ("A" <= c <= "Z") or c in "-.*" for c in self._seq):
            raise CheckError("InvalidSequence")

    def id(self) -> str:
        return self._id

    def desc(self) -> Optional[str]:
        return self._desc

    def seq(self) -> bytes:
        return self._seq.encode("utf-8")

    def clear(self) -> None:
        self._id, self._desc, self._seq = "", None, ""

    def __str__(self) -> str:  # fasta.rs:1037-1080 (Display)
        head = ">" + self._id + ((" " + self._desc) if self._desc is not None else "")
        return head + "\n" + self._seq + "\n"


def _text_stream(src):
    if isinstance(src, (bytes, bytearray, memoryview)):
        return io.StringIO(bytes(src).decode("utf-8"), newline="")
    if isinstance(src, str):
        return io.StringIO(src, newline="")
    if isinstance(src, io.TextIOBase):
        return src
    return io.TextIOWrapper(src, encoding="utf-8", newline="")


class Reader:
    """fasta.rs:174-360.  `Reader.new(bytes | str | binary/text file object)`, `Reader.from_file(path)`."""

    def __init__(self, stream):
        self._stream = stream
        self._line = ""

    @staticmethod
    def new(reader) -> "Reader":
        return Reader(_text_stream(reader))

    @staticmethod
    def from_file(path) -> "Reader":
        return Reader(open(path, "r", encoding="utf-8", newline=""))

    def read(self, record: Record) -> None:
        """fasta.rs:334-359: fills `record`; an empty record means the input is exhausted.  OSError where the
        reference returns io::Error ("Expected > at record start.")."""
        record.clear()
        if not self._line:
            self._line = self._stream.readline()
            if not self._line:
                return
        if not self._line.startswith(">"):
            raise OSError("Expected > at record start.")
        head = self._line[1:].rstrip()
        # splitn(2, char::is_whitespace): the id ends at the first whitespace character, the rest is the description
        cut = next((i for i, c in enumerate(head) if c.isspace()), None)
        if cut is None:
            record._id, record._desc = head, None
        else:
            record._id, record._desc = head[:cut], head[cut + 1:]
        parts = []
        while True:
            self._line = self._stream.readline()
            if not self._line or self._line.startswith(">"):
                break
            parts.append(self._line.rstrip())
        record._seq = "".join(parts)

    def records(self) -> Iterator[Record]:
        """fasta.rs:287, 1082-1110"""
        while True:
            r = Record()
            self.read(r)
            if r.is_empty():
                return
            yield r
