"""``metacov`` command line on the B200 path.

Keeps the commands, options and CSV formats of reference metacov/cli.py
(`pileup` cli.py:35-109, `scan` cli.py:112-285) so that existing invocations
keep working; the work behind them runs on the GPU.  What is not carried over
is out of the hot-path scope and fails with a clear message instead of being
silently approximated: FASTQ input and `-b/-M` of `scan`, and the `simulate`
command.  `pileup -k` (the `experimental` coverage model,
reference metacov/pileup.py:38-173) runs on the GPU like `classic`.
"""
import csv
import logging
import sys

import click

from . import pileup as _pileup
from . import scan as _scan
from . import util
from .alignmentfile import AlignmentFile
from .fasta import FastaFile

logging.basicConfig(level=logging.INFO, format="[%(relativeCreated)6.1f %(funcName)s]  %(message)s",
                    datefmt="%I:%M:%S")
log = logging.getLogger(__name__)


@click.group()
def main():
    """
    MetaCov estimates abundance values from the stacking depth of
    reads mapped to a reference.
    """


@main.command()
@click.option("--bamfile", "-b", type=click.File("rb"), required=True,
              help="Input BAM file. Must be sorted and indexed.")
@click.option("--reference-fasta", "-f", type=click.File("rb"))
@click.option("--regionfile-blast7", "-rb", type=click.File("r"), help="Input Region file in BLAST7 format")
@click.option("--regionfile-csv", "-rc", type=click.File("r"), help="Input Region file in CSV format")
@click.option("--kmer-histogram", "-k", type=click.File("r"), help="Kmer Histogram produced with metacov scan")
@click.option("--kmer-length", "-K", type=int, default=7, help="Length of k-mer")
@click.option("--outfile", "-o", type=click.File("w"), default="-", help="Output CSV (default STDOUT)")
@click.option("--bedgraph", "-bg", type=click.File("w"), metavar="FILE",
              help="Also write the per-base depth as bedGraph (runs of equal depth; zero-depth runs omitted). "
                   "Not in the reference: additive.")
@click.option("--window", "-w", type=click.IntRange(1), metavar="N",
              help="Also write the mean depth of fixed windows of N bp to --window-out. Not in the reference: additive.")
@click.option("--window-out", "-wo", type=click.File("w"), metavar="FILE", help="Output CSV of --window (sacc,start,end,avg)")
@click.option("--bam-decode", type=click.Choice(["host", "gpu", "gpu-stream", "auto"]), default="auto", show_default=True,
              help="Where the BAM file is inflated and parsed: host threads (zlib), the GPU (whole file / streamed in chunks), or "
                   "auto = the GPU, streamed when the file does not fit it. Not in the reference: additive.")
def pileup(bamfile, reference_fasta, regionfile_blast7, regionfile_csv, kmer_histogram, kmer_length, outfile,
           bedgraph=None, window=None, window_out=None, bam_decode="auto"):
    """
    Compute fold coverage values
    """
    if (window is None) != (window_out is None):
        raise click.UsageError("--window and --window-out go together")
    bam = AlignmentFile(bamfile.name, decode=bam_decode)
    fasta = FastaFile(reference_fasta.name) if reference_fasta else None          # cli.py:59
    regions = util.make_region_iterator(regionfile_blast7, regionfile_csv, bam)
    try:
        mapped, unmapped = bam.mapped, bam.unmapped
        log.info("Number of reads:\n  total:    {total}\n  mapped:   {mapped} ({mappct}%)\n  unmapped: {unmapped}\n"
                 "".format(mapped=mapped, unmapped=unmapped, total=mapped + unmapped,
                           mappct=mapped / (mapped + unmapped) * 100))
    except (AttributeError, ValueError, ZeroDivisionError):
        log.error("BAM file not indexed!?")

    name2ref = {word.split()[0]: word for word in bam.references}
    # The reference piles every region up from scratch (cli.py:85-95); here the depth of the
    # whole file exists once on the GPU and all regions are answered by one kernel launch.
    hits = list(regions)
    refs, starts, ends = [], [], []
    for hit in hits:
        refs.append(name2ref[hit.sacc])
        # blast uses end<start for the reverse strand; the sorted pair is used verbatim as
        # 0-based half-open (cli.py:89)
        start, end = sorted((int(hit.sstart), int(hit.send)))
        starts.append(start)
        ends.append(end)
    # cli.py:81 -- the k-mer length option is not passed on to the loader (its default 7 applies)
    k_cor = _pileup.load_kmerhist(kmer_histogram) if kmer_histogram else None
    with click.progressbar(file=sys.stderr, length=0, label="Calculating coverages"):
        results = _pileup.classic_many(bam, refs, starts, ends) if hits else []
        if k_cor is not None and hits:                                            # cli.py:93-95
            for result, extra in zip(results, _pileup.experimental_many(bam, k_cor, kmer_length, fasta, refs, starts, ends)):
                result.update(extra)
    writer = None
    for hit, result in zip(hits, results):
        if writer is None:
            writer = csv.DictWriter(outfile, fieldnames=["sacc", "start", "end"] + sorted(result.keys()))
            writer.writeheader()
        result.update({"sacc": hit.sacc, "start": hit.sstart, "end": hit.send})
        writer.writerow(result)
    if bedgraph is not None:
        write_bedgraph(bam, bedgraph)
    if window is not None:
        write_windows(bam, window, window_out)
    bam.close()


def write_bedgraph(bam, out):
    """Per-base depth as bedGraph lines `name<TAB>start<TAB>end<TAB>depth` (0-based half-open), one per
    run of equal non-zero depth; the runs come from the GPU (mcov_depth_runs)."""
    names = [ref.split()[0] for ref in bam.references]
    runs = bam.coverage_engine().depth_runs(skip_zero=True)
    for t, s, e, d in zip(runs["tid"].tolist(), runs["start"].tolist(), runs["end"].tolist(), runs["depth"].tolist()):
        out.write("%s\t%d\t%d\t%d\n" % (names[t], s, e, d))


def write_windows(bam, window, out):
    """Mean depth of fixed windows (the last window of a contig may be shorter): CSV sacc,start,end,avg."""
    means = bam.coverage_engine().window_means(window)
    w = csv.writer(out)
    w.writerow(["sacc", "start", "end", "avg"])
    k = 0
    for ref, ln in zip(bam.references, bam.lengths):
        for p in range(0, ln, window):
            w.writerow([ref.split()[0], p, min(p + window, ln), round(float(means[k]), 2)])
            k += 1


@main.command()
@click.option("--readfile-type", "-t", type=click.Choice(["bam", "fq"]),
              help="Override filename based detection of readfile type.")
@click.option("--max-reads", "-m", type=click.IntRange(1), metavar="N", help="Only consider the first N reads.")
@click.option("--group-by", "-g", type=click.Choice(list(_scan.Flags.keys())), multiple=True, metavar="FLAG",
              help="Group output by BAM flag. May be specified multiple times. "
                   "FLAG can be one of {}".format(list(_scan.Flags.keys())))
@click.option("--reference-fasta", "-f", type=click.File("rb"), metavar="FILE",
              help="Fasta file reads where mapped to. Required for boffset. File may be bgzip'ed, but not gzip'ed.")
@click.option("--out-basehist", "-b", type=click.File("w", lazy=False), metavar="FILE",
              help="Compute histogram of base counts by position in read.")
@click.option("--boffset", "-bo", type=click.IntRange(0, 200), default=0, metavar="N", show_default=True,
              help="Include N bases prior to read start in base histogram. Requires reference fasta file.")
@click.option("--out-kmerhist", "-o", type=click.File("w", lazy=False), metavar="FILE",
              help="Compute histogram of kmers in reads")
@click.option("-k", type=click.IntRange(2, 12), default=7, metavar="N", show_default=True, help="Length of kmers")
@click.option("--step", "-s", type=click.IntRange(1, 100), default=7, metavar="N", show_default=True,
              help="Step between sampled kmers")
@click.option("--offset", "-O", type=click.IntRange(-100, 100), default=0, metavar="N", show_default=True,
              help="Offset of first sampled kmer")
@click.option("--number", "-n", type=click.IntRange(1, 100), default=8, metavar="N", show_default=True,
              help="Numer of sampled kmers")
@click.option("--out-mirrorhist", "-M", type=click.File("w", lazy=False), metavar="FILE",
              help="Compute histogram of mismatches against palindrome")
@click.option("--mirror-offset", "-MO", type=click.IntRange(-100, 100), default=4, metavar="N", show_default=True,
              help="Offset of palindrome center from read start")
@click.option("--mirror-length", "-Ml", type=click.IntRange(1, 50), default=10, metavar="N", show_default=True,
              help="Palindrome length is 2N+1")
@click.option("--out-isizehist", "-I", type=click.File("w", lazy=False), metavar="FILE",
              help="Compute histogram of insert sizes.")
@click.option("--bam-decode", type=click.Choice(["host", "gpu", "gpu-stream", "auto"]), default="auto", show_default=True,
              help="Where the BAM file is inflated and parsed (see `pileup`); with the GPU the accumulators run on the "
                   "device-resident columns. Not in the reference: additive.")
@click.argument("readfile", nargs=-1, required=True)
def scan(readfile, readfile_type, out_basehist, boffset, out_kmerhist, k, number, step, offset, max_reads,
         reference_fasta, out_mirrorhist, mirror_offset, mirror_length, out_isizehist, group_by, bam_decode="auto"):
    """
    Gather read statistics
    """
    if not readfile_type:
        for ext, ft in ((".bam", "bam"), (".sam", "bam"), (".fq", "fq"), (".fq.gz", "fq"), (".fastq", "fq"),
                        (".fastq.gz", "fq")):
            if readfile[0].endswith(ext):
                readfile_type = ft
                break
    if not readfile_type:
        raise click.UsageError("Couldn't guess input format. Please supply -t")
    if readfile_type != "fq" and len(readfile) > 1:
        raise click.UsageError("Multiple input files only supported for fastq")
    if len(readfile) > 2:
        raise click.UsageError("At most two fastq files allowed (fwd and rev)")
    if readfile_type != "bam" and reference_fasta:
        raise click.UsageError("Reference fasta can only be used with mapped (bam/sam) reads")
    if readfile_type != "bam":
        raise click.UsageError("FASTQ input is outside the GPU hot path of this build (unmapped reads carry "
                               "no coverage); use the reference implementation for it")
    for opt, name in ((out_basehist, "-b/--out-basehist"), (out_mirrorhist, "-M/--out-mirrorhist")):
        if opt:
            raise click.UsageError(name + " needs the reference FASTA and is outside the GPU hot path of this build")

    update_every = 100000
    if max_reads and max_reads > 0 and max_reads / 100 < update_every:
        update_every = int(max_reads + 50 / 100)        # sic (reference cli.py:205-206)

    infile = AlignmentFile(readfile[0], decode=bam_decode)
    log.info("mapped = {}, unmapped = {}, total = {}".format(infile.mapped, infile.unmapped,
                                                             infile.mapped + infile.unmapped))
    length = infile.mapped + infile.unmapped
    if max_reads and 0 < max_reads < length:
        length = max_reads

    counters = []
    if out_kmerhist:
        counters.append(_scan.KmerHist(k, number, step, offset))
    if out_isizehist:
        counters.append(_scan.IsizeHist())
    counters = _scan.ByFlag(counters, [_scan.Flags[flag] for flag in group_by])

    with click.progressbar(length=length, label="Scanning reads", show_pos=True) as bar:
        with infile:
            nreads = _scan.scan_reads(infile, None, counters, update_every, lambda: bar.update(update_every),
                                      max_reads)
        bar.update(update_every)
    log.info("Processed {} reads".format(nreads))

    n = 0
    if out_kmerhist:
        csv.writer(out_kmerhist).writerows(counters.get_rows(n))
        n += 1       # (the reference forgets this increment, cli.py:273-280; harmless there without -M)
    if out_isizehist:
        csv.writer(out_isizehist).writerows(counters.get_rows(n))
        n += 1


if __name__ == "__main__":
    main()
