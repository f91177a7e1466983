"""Helpers shared by the GPU parity tests."""
import numpy as np

import oracle_lib
from bio_ik_b200 import workloads
from bio_ik_b200.solver import IKSolver

MODE_NAMES = {"q": "bio2_memetic", "l": "bio2_memetic_l", 0: "bio2"}
TRACE_KEYS = ("genes", "gradients", "species_fitness", "solutions", "fitness")


def make_solver(w, population, mode="q", generations=8, random_seed=1):
    return IKSolver(w.robot, mode=MODE_NAMES[mode], population=population, generations=generations, random_seed=random_seed, device=0).initialize(w.problem)


def assert_bit_equal(got, ref, keys=TRACE_KEYS, what=""):
    for k in keys:
        a, b = np.asarray(got[k]), np.asarray(ref[k])
        if not np.array_equal(a, b):
            d = np.abs(a.astype(np.float64) - b.astype(np.float64))
            raise AssertionError(f"{what} {k}: not bit-identical to the oracle; max abs diff {np.nanmax(d):.3e} at {np.unravel_index(np.nanargmax(d), d.shape)}")
