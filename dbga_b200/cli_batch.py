"""`dbga-batch`: the `dbga` command line for a file of many pairs (records a1, b1, a2, b2, ...).

Same scoring flags and defaults as the reference's CLI (src/dbga/cli.py:11-86); every pair gets what
`dbga -i pair.fasta -o out.fasta -k K -m dna` writes for it (cli.py:131-136: the two gapped rows as FASTA
records wrapped at 60 columns), pairs the reference rejects (cycles, k larger than a sequence) are
reported and skipped.  The file is parsed straight into pinned memory and 2-bit words, aligned on
`--gpus` GPUs (pairs sharded by index, no collective), and written from the compact result form."""
from __future__ import annotations

import time
from pathlib import Path

import click
import numpy as np


@click.command()
@click.option("--infile", "-i", type=click.Path(exists=True), required=True, help="FASTA file of pairs: records a1, b1, a2, b2, ...")
@click.option("--outfile", "-o", type=click.Path(exists=False), required=True, help="Output aligned sequences file")
@click.option("-k", type=int, required=True, help="The kmer size for constructing de Bruijn graph")
@click.option("--moltype", "-m", type=str, default="dna", help="The input sequence molecular type")
@click.option("--match", default=10, type=int, help="score for two matching nucleotide")
@click.option("--transition", default=-1, type=int, help="cost for DNA transition mutation")
@click.option("--transversion", default=-8, type=int, help="cost for DNA transversion mutation")
@click.option("-d", default=10, type=int, help="costs for opening a gap")
@click.option("-e", default=2, type=int, help="costs for extending a gap")
@click.option("--gpus", default=1, type=int, help="GPUs of this box to shard the pairs over")
def main(infile, outfile, k, moltype, match, transition, transversion, d, e, gpus):
    from . import _lib
    from .batch import Aligner, MultiAligner
    from .fasta import FastaBatch
    infile_path, outfile_path = Path(infile), Path(outfile)
    if infile_path.suffix.lower() != ".fasta" or outfile_path.suffix.lower() != ".fasta":
        click.secho("Expect input and output files to be `fasta` format!", err=True, fg="red")
        return
    t0 = time.perf_counter()
    fa = FastaBatch(infile_path, want_packed=True)
    if fa.num_records < 2 or fa.num_records % 2:
        click.secho("A pair file needs an even, non-zero number of sequences!", err=True, fg="red")
        return
    t1 = time.perf_counter()
    P = fa.num_pairs
    click.secho(f"Number of pairs: {P}, running de Bruijn pairwise alignment on {gpus} GPU(s)", fg="green")
    prm = Aligner.params(k, match, transition, transversion, d, e, flags=_lib.FLAG_COMPACT_ROWS)
    packed = fa.bad_records == 0            # non-ACGT characters cannot be packed: the ASCII path reports them per pair
    offs = np.ascontiguousarray(fa.offsets, np.int64)
    written, status = 0, []
    if gpus > 1:
        m = MultiAligner(list(range(gpus)))
        res = m.align_buffers(None if packed else fa.seqs_ptr, offs, prm, packed=fa.packed_ptr if packed else None, unpack=False)
        t2 = time.perf_counter()
        for s in range(res.num_shards):
            written += fa.write(outfile_path, res.shard[s], first_record=2 * int(res.pair_lo[s]), append=s > 0)
            n = int(res.shard[s].num_pairs)
            status.append(np.ctypeslib.as_array(res.shard[s].status, shape=(n,)).copy() if n else np.zeros(0, np.int32))
        m.close()
    else:
        al = Aligner(0)
        raw = al.align_buffers(fa.packed_ptr if packed else fa.seqs_ptr, offs, prm, unpack=False, packed=packed)
        t2 = time.perf_counter()
        written = fa.write(outfile_path, raw)
        status.append(np.ctypeslib.as_array(raw.status, shape=(P,)).copy())
        al.close()
    t3 = time.perf_counter()
    status = np.concatenate(status)
    for code, what in ((_lib.CYCLE, "cycles in the de Bruijn graph (usually caused by small kmer sizes)"),
                       (_lib.K_INVALID, "invalid k size for kmers"), (_lib.BAD_CHAR, "characters outside ACGT"),
                       (_lib.TOO_LARGE, "a workspace or numeric limit")):
        c = int((status == code).sum())
        if c:
            click.secho(f"{c} pair(s) not aligned: {what}", err=True, fg="red")
    click.secho(f"Alignment task finishes within {round(t3 - t0, 2)} seconds! "
                f"(read {t1 - t0:.3f} s, align {t2 - t1:.3f} s, write {t3 - t2:.3f} s; {written} pairs written)", fg="green")
    click.secho(f"Alignment results saved in {outfile_path.absolute()}!", fg="green")


if __name__ == "__main__":
    main()
