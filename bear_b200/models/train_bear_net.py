"""Train and evaluate AR or BEAR models from a config file. Usage:

``python -m bear_b200.models.train_bear_net config.cfg``

Drop-in for the reference's ``bear_model/models/train_bear_net.py`` (same config keys, result keys and
output files); example configs are in ``bear_b200/models/config_files``.
"""
import argparse
import configparser

from bear_b200 import bear_net
from bear_b200.models import _script


def main(config):
    return _script.run(config, bear_net, is_ref=False)


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument('configPath')
    args = parser.parse_args()
    config = configparser.ConfigParser()
    config.read(args.configPath)
    main(config)
