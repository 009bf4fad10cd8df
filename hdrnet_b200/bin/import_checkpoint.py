#!/usr/bin/env python
"""Convert a TensorFlow training directory of the reference (``checkpoint`` state file +
``model.ckpt-N.{index,data-*,meta}``, what hdrnet/bin/run.py:136-142 restores) into the
``weights.npz`` + ``params.json`` pair ``hdrnet_b200.bin.run`` loads:

    python -m hdrnet_b200.bin.import_checkpoint <tf_checkpoint_dir> <out_dir> [--params params.json]
                                                [--legacy] [--verify] [--guide_bins]

``--legacy`` applies the old-variable-name map of scripts/upgrade.py; ``--guide_bins`` also writes
the raw float32 guide dumps of hdrnet/bin/freeze_graph.py:105-185 next to the weights.
"""
from __future__ import annotations

import argparse
import json
import logging

from hdrnet_b200 import checkpoint
from hdrnet_b200.bin.run import save_checkpoint

log = logging.getLogger("import_checkpoint")


def main(args):
    params, weights = checkpoint.import_checkpoint(args.checkpoint_dir, verify=args.verify, legacy=args.legacy)
    if args.params:
        with open(args.params) as f:
            params = json.load(f)
    if params is None:
        raise SystemExit("no .meta file next to the checkpoint: pass the model parameters with --params")
    if not weights:
        raise SystemExit("the checkpoint holds no variables under 'inference/' (try --legacy)")
    save_checkpoint(args.out_dir, params, weights)
    log.info("wrote %d variables to %s", len(weights), args.out_dir)
    if args.guide_bins:
        for p in checkpoint.export_guide_bins(weights, params["model_name"], args.out_dir):
            log.info("wrote %s", p)


if __name__ == "__main__":
    logging.basicConfig(level=logging.INFO)
    ap = argparse.ArgumentParser()
    ap.add_argument("checkpoint_dir")
    ap.add_argument("out_dir")
    ap.add_argument("--params", default=None, help="model_params as JSON (when there is no .meta file)")
    ap.add_argument("--legacy", action="store_true", help="old variable names (scripts/upgrade.py)")
    ap.add_argument("--verify", action="store_true", help="check every CRC-32C (slow, pure Python)")
    ap.add_argument("--guide_bins", action="store_true", help="also write freeze_graph.py's guide .bin dumps")
    main(ap.parse_args())
