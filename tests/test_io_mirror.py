"""The I/O step either side of the path (rust_bio_b200.io): the reference's doc examples for the FASTA / FASTQ
readers (src/io/fasta.rs:17-60, 318-330, 958-970; src/io/fastq.rs:205-217, 388-410) and the batch packing."""
import io

import numpy as np
import pytest

from rust_bio_b200.io import fasta, fastq, records_to_batch


def test_fasta_reader_doc_examples():
    r = fasta.Reader.new(b">id desc\nAAAA\n")
    rec = fasta.Record.new()
    r.read(rec)
    assert (rec.id(), rec.desc(), rec.seq()) == ("id", "desc", b"AAAA")
    r.read(rec)
    assert rec.is_empty()
    recs = list(fasta.Reader.new(io.BytesIO(b">a first one\nACGT\nAC\n>b\nTT\r\nGG\r\n")).records())
    assert [(x.id(), x.desc(), x.seq()) for x in recs] == [("a", "first one", b"ACGTAC"), ("b", None, b"TTGG")]
    for x in recs:
        x.check()
    assert str(fasta.Record.with_attrs("read1", "sampleid=foobar", b"ACGT")) == ">read1 sampleid=foobar\nACGT\n"
    with pytest.raises(OSError, match="Expected > at record start"):
        list(fasta.Reader.new(b"ACGT\n").records())
    with pytest.raises(fasta.CheckError, match="InvalidSequence"):
        fasta.Record.with_attrs("x", None, b"AC1T").check()
    with pytest.raises(fasta.CheckError, match="EmptyId"):
        fasta.Record.with_attrs("", None, b"ACGT").check()


def test_fastq_reader_doc_examples():
    recs = list(fastq.Reader.new(b"@id description\nACGT\n+\n!!!!\n").records())
    assert len(recs) == 1 and (recs[0].id(), recs[0].desc(), recs[0].seq(), recs[0].qual()) == ("id", "description", b"ACGT", b"!!!!")
    recs[0].check()
    multi = list(fastq.Reader.new(b"@r1\nACGT\nGGCC\n+\n!!!!\n####\n@r2 x y\nTT\n+r2\nII\n").records())
    assert [(x.id(), x.desc(), x.seq(), x.qual()) for x in multi] == [("r1", None, b"ACGTGGCC", b"!!!!####"), ("r2", "x y", b"TT", b"II")]
    with pytest.raises(fastq.ReadError, match="MissingAt"):
        list(fastq.Reader.new(b"id\nACGT\n+\n!!!!\n").records())
    with pytest.raises(fastq.ReadError, match="IncompleteRecord"):
        list(fastq.Reader.new(b"@id\nACGT\n+\n").records())
    with pytest.raises(fasta.CheckError, match="UnequalLength"):
        fastq.Record.with_attrs("a", None, b"ACGT", b"!!").check()


def test_records_to_batch_layout():
    reads = list(fastq.Reader.new(b"@r1\nACGTAC\n+\n!!!!!!\n@r2\nTTG\n+\n!!!\n").records())
    refs = list(fasta.Reader.new(b">w1\nGGACGTACGG\n>w2\nATTGA\n").records())
    (blob, x_off, x_len, y_off, y_len), keep = records_to_batch(reads, refs)
    assert keep is None and list(x_len) == [6, 3] and list(y_len) == [10, 5]
    assert all(int(o) % 16 == 0 for o in list(x_off) + list(y_off))
    assert bytes(blob[int(x_off[1]):int(x_off[1]) + 3]) == b"TTG" and bytes(blob[int(y_off[0]):int(y_off[0]) + 10]) == b"GGACGTACGG"
    with pytest.raises(ValueError):
        records_to_batch(reads, refs[:1])
