#!/usr/bin/env python
"""Exports the reference's pickled benchmark histories (``OrderedDict{ConfigSpace.Configuration -> perf}``,
reference tools/offline_benchmark.py:154-176) as JSON lines for ``tlbo_b200.config_space.space.load_history``.
Run it WHERE ConfigSpace 0.4.12 and the pickles live (it imports nothing from this repository):

    python export_history.py data/hpo_data/<task>.pkl > <task>.jsonl
"""
import json
import pickle
import sys


def main():
    with open(sys.argv[1], 'rb') as f:
        hist = pickle.load(f)
    for config, perf in hist.items():
        values = {k: (v.item() if hasattr(v, 'item') else v) for k, v in config.get_dictionary().items()}
        print(json.dumps({'config': values, 'perf': perf if not hasattr(perf, 'tolist') else perf.tolist()}))


if __name__ == '__main__':
    main()
