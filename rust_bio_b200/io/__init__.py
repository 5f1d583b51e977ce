"""The callers' I/O step either side of the alignment path: `bio::io::fasta` /
`bio::io::fastq` readers (reference src/io/fasta.rs, src/io/fastq.rs) restated as host code, and the packing of
their records into the batch buffers the engine takes (pinned when torch is given).  Pure host work; no kernels.
"""
from . import fasta, fastq  # noqa: F401
from .batch import pairs_from_records, records_to_batch  # noqa: F401
