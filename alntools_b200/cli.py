"""`python -m alntools_b200 bam2ec|bam2emase ...` with the reference's options (alntools/cli.py:43-89)."""
import click

from . import methods, utils


@click.group(context_settings=dict(help_option_names=['-h', '--help']))
def cli():
    """alntools bam2ec / bam2emase on a B200"""


def _common(fn):
    opts = [
        click.option('-c', '--chunks', default=0, help="number of chunks to process"),
        click.option('-d', '--directory', type=click.Path(exists=True, resolve_path=True, file_okay=False,
                                                          dir_okay=True, writable=True), help="temp directory"),
        click.option('--multisample', is_flag=True),
        click.option('-p', '--number-processes', default=-1, help="number of processes"),
        click.option('--rangefile', type=click.Path(exists=False, resolve_path=True, file_okay=True,
                                                    dir_okay=False, writable=True), help="range file"),
        click.option('-t', '--targets', metavar='FILE', type=click.Path(exists=True, resolve_path=True,
                                                                       file_okay=True, dir_okay=False),
                     help="target file"),
        click.option('-v', '--verbose', count=True, help='enables verbose mode'),
    ]
    for opt in reversed(opts):
        fn = opt(fn)
    return fn


@cli.command('bam2ec', options_metavar='<options>', short_help='convert a BAM file to EC')
@click.argument('bam_file', metavar='bam_file', type=click.Path(exists=True, resolve_path=True, dir_okay=True))
@click.argument('ec_file', metavar='ec_file', type=click.Path(resolve_path=True, dir_okay=False, writable=True))
@click.option('-m', '--mincount', default=1000, help="minimum count")
@click.option('-s', '--sample', help="sample identifier")
@_common
def bam2ec(bam_file, ec_file, chunks, directory, mincount, multisample, number_processes, rangefile, sample,
           targets, verbose):
    """Convert a BAM file (bam_file) to a binary EC file (ec_file)"""
    utils.configure_logging(verbose)
    if multisample:
        if sample:
            print('-s, --sample should NOT be specified with --multisample')
            return
        methods.bam2ec_multisample(bam_file, ec_file, chunks, mincount, directory, number_processes, rangefile,
                                   targets)
    else:
        methods.bam2ec(bam_file, ec_file, chunks, directory, number_processes, rangefile, sample, targets)


@cli.command('bam2emase', options_metavar='<options>', short_help='convert a BAM file to EMASE format')
@click.argument('bam_file', metavar='bam_file', type=click.Path(exists=True, resolve_path=True, dir_okay=True))
@click.argument('emase_file', metavar='emase_file', type=click.Path(resolve_path=True, dir_okay=False,
                                                                    writable=True))
@click.option('-m', '--mincount', default=2000, help="minimum count")
@_common
def bam2emase(bam_file, emase_file, chunks, directory, mincount, multisample, number_processes, rangefile,
              targets, verbose):
    """Convert a BAM file (bam_file) to an EMASE file (emase_file)"""
    utils.configure_logging(verbose)
    if multisample:
        methods.bam2emase_multisample(bam_file, emase_file, chunks, mincount, directory, number_processes,
                                      rangefile, targets)
    else:
        methods.bam2emase(bam_file, emase_file, chunks, directory, number_processes, rangefile, targets)


@cli.command('ec2emase', options_metavar='<options>', short_help='convert a binary EC file to EMASE format')
@click.argument('ec_file', metavar='ec_file', type=click.Path(exists=True, resolve_path=True, dir_okay=False))
@click.argument('emase_file', metavar='emase_file', type=click.Path(resolve_path=True, dir_okay=False))
@click.option('-v', '--verbose', count=True, help='enables verbose mode')
def ec2emase(ec_file, emase_file, verbose):
    """Convert a binary EC format file (ec_file) to EMASE format (emase_file)"""
    utils.configure_logging(verbose)
    methods.ec2emase(ec_file, emase_file)
