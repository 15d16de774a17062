"""Train and evaluate reference-based AR or BEAR models from a config file. Usage:

``python -m bear_b200.models.train_bear_ref config.cfg``

Drop-in for the reference's ``bear_model/models/train_bear_ref.py``; adds ``error_rate`` and
``stop_rate`` to ``[results]`` (models/train_bear_ref.py:143-147).  Example configs:
``config_files/bear_stop_bear.cfg`` and ``bear_stop_ar.cfg``.
"""
import argparse
import configparser

from bear_b200 import bear_ref
from bear_b200.models import _script


def main(config):
    return _script.run(config, bear_ref, is_ref=True)


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument('configPath')
    args = parser.parse_args()
    config = configparser.ConfigParser()
    config.read(args.configPath)
    main(config)
