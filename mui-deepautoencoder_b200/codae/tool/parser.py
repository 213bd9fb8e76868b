import argparse


def parse():
    """Stale helper kept for import parity (codae/tool/parser.py:5-27 is dead code in the reference: both
    scripts shadow it with their own parse())."""
    parser = argparse.ArgumentParser(description='Train denoising autoencoder.')
    parser.add_argument('--config', type=str, default=None)
    parser.add_argument('--debug', type=bool, default=False)
    return parser.parse_known_args()[0]
