"""Shared helpers for the parity tests (test infrastructure)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "reach_golden.npz")


def load_golden():
    return np.load(GOLDEN)


def golden_case(g, name):
    prefix = name + "__"
    return {k[len(prefix):]: g[k] for k in g.files if k.startswith(prefix)}


def oracle_chain():
    from oracle.reach_oracle import OracleChain
    from pioneer_b200.urdf import flatten_urdf
    return OracleChain.from_model(flatten_urdf())
