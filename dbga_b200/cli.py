"""`dbga` command line (same flags, defaults and messages as the reference's
src/dbga/cli.py:10-160).  Two sequences -> pairwise de Bruijn alignment on the GPU.  More
than two sequences is the reference's MSA path (cogent3 progressive pair-HMM), which is out
of this package's scope and reported as such."""
from __future__ import annotations

from pathlib import Path
from time import time

import click

from .debruijn_pairwise import DeBruijnPairwise
from .utils import load_sequences


@click.command()
@click.option("--infile", "-i", type=click.Path(exists=True), required=True, help="Input original sequences file")
@click.option("--outfile", "-o", type=click.Path(exists=False), required=True, help="Output aligned sequences file")
@click.option("-k", type=int, required=True, help="The kmer size for constructing de Bruijn graph")
@click.option("--moltype", "-m", type=str, required=True, help="The input sequence molecular type")
@click.option("--match", default=10, type=int, required=False, help="score for two matching nucleotide (pairwise only)")
@click.option("--transition", default=-1, type=int, required=False, help="cost for DNA transition mutation (pairwise only)")
@click.option("--transversion", default=-8, type=int, required=False, help="cost for DNA transversion mutation (pairwise only)")
@click.option("-d", default=10, type=int, required=False, help="costs for opening a gap (pairwise only)")
@click.option("-e", default=2, type=int, required=False, help="costs for extending a gap (pairwise only)")
@click.option("--model", type=str, required=False, default="F81", help="The model for multiple sequence alignment (MSA only)")
@click.option("--indel_rate", type=float, required=False, default=0.01, help="One parameter for the progressive pair-HMM (MSA only)")
@click.option("--indel_length", type=float, required=False, default=0.01, help="One parameter for the progressive pair-HMM (MSA only)")
def main(infile, outfile, k, moltype, match, transition, transversion, d, e, model, indel_rate, indel_length):
    infile_path, outfile_path = Path(infile), Path(outfile)
    if infile_path.suffix.lower() != ".fasta" or outfile_path.suffix.lower() != ".fasta":
        click.secho("Expect input and output files to be `fasta` format!", err=True, fg="red")
        return
    sequence_collection = load_sequences(data=infile_path, moltype=moltype)
    sequence_num = sequence_collection.num_seqs
    start_time = time()
    if sequence_num < 2:
        click.secho("Too few sequences for alignment, should be at least 2!", err=True, fg="red")
        return
    if sequence_num == 2:
        click.secho(f"Number of sequences: {sequence_num}, running de Bruijn pairwise alignment", fg="green")
        dbg = DeBruijnPairwise(data=sequence_collection, k=k, moltype=moltype)
        aln = dbg.alignment(match=match, transition=transition, transversion=transversion, d=d, e=e)
        if isinstance(aln, dict):
            # the reference returns a bare dict when the sequences share no k-mer
            # (debruijn_pairwise.py:453-457) and then fails on aln.to_fasta(); write FASTA.
            from .seqcoll import make_aligned_seqs
            aln = make_aligned_seqs(aln, moltype=moltype)
        with open(outfile, "w") as f:
            f.write(aln.to_fasta())
    else:
        click.secho(
            f"Number of sequences: {sequence_num}: multiple sequence alignment is outside the scope of dbga_b200 "
            "(the reference delegates it to cogent3's progressive pair-HMM)", err=True, fg="red")
        return
    click.secho(f"Alignment task finishes within {round(time() - start_time, 2)} seconds!", fg="green")
    click.secho(f"Alignment results saved in {outfile_path.absolute()}!", fg="green")


if __name__ == "__main__":
    main()
