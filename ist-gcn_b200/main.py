#!/usr/bin/env python
"""Entry point with the reference's command line (main.py:1-33):

    python main.py recognition -c config/st_gcn/<dataset>/train.yaml [--key value ...]
    torchrun --nproc-per-node 8 main.py recognition -c ...          # one process per GPU

Run it from this directory (or with it on PYTHONPATH) so that ``net``, ``feeder`` and ``processor``
resolve to the drop-in packages next to this file."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    from processor.recognition import REC_Processor
    processors = {'recognition': REC_Processor}
    parser = argparse.ArgumentParser(description='Processor collection')
    subparsers = parser.add_subparsers(dest='processor')
    for k, p in processors.items():
        subparsers.add_parser(k, parents=[p.get_parser()])
    arg = parser.parse_args()
    if arg.processor is None:
        parser.print_help()
        return 2
    p = processors[arg.processor](sys.argv[2:])
    p.start()
    return 0


if __name__ == '__main__':
    sys.exit(main())
