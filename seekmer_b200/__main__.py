#!/usr/bin/env python3
"""`seekmer` command line (`__main__.py:13-71`): the `infer` and `impute` sub-commands are
served by this package; `index` stays with the reference (out of scope)."""
import argparse
import sys

from . import impute
from . import infer
from ._log import StderrHandler


def main(argv=None):
    parser = argparse.ArgumentParser(prog='seekmer', description='A fast RNA-seq tool (B200 infer path)',
                                     add_help=True)
    parser.add_argument('-v', '--version', action='version', version='Seekmer 2019.0.0 (seekmer_b200)')
    parser.add_argument('--debug', action='store_true', help='enable debugging messages')
    subparsers = parser.add_subparsers(title='subcommand', dest='subcommand')
    infer.add_subcommand_parser(subparsers)
    impute.add_subcommand_parser(subparsers)
    opts = vars(parser.parse_args(argv))
    handler = StderrHandler(level='DEBUG' if opts['debug'] else 'INFO')
    with handler.applicationbound():
        if opts['subcommand'] == 'infer':
            infer.run(**opts)
        elif opts['subcommand'] == 'impute':
            impute.run(**opts)
        else:
            parser.print_help()
    return 0


if __name__ == '__main__':
    sys.exit(main())
