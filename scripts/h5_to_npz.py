#!/usr/bin/env python
"""Convert the reference's HDF5 rollouts (scripts/collect_data.py:20-44,137-170) to .npz shards that
swarm_ode_b200.dataset.load_episodes reads without h5py.  Run wherever h5py is installed:

    python scripts/h5_to_npz.py warehouse_data_<env>_seed0.h5 [...]      ->  warehouse_data_<env>_seed0.npz

Layout read: episode_%06d/metadata attrs {num_agvs, num_pickers}, episode_%06d/steps/step_%06d/observations [n, D]
(ragged rows are zero-padded to the widest, as collect_data.py:99-118 does).  Written per episode:
episode_%06d/observations [n_steps, n, D] float32, episode_%06d/num_agvs, episode_%06d/num_pickers."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from swarm_ode_b200.dataset import load_episodes
    for path in sys.argv[1:]:
        eps = load_episodes(path)      # the .h5 branch (needs h5py)
        out = {}
        for i, ep in enumerate(eps):
            out[f"episode_{i:06d}/observations"] = ep.observations
            out[f"episode_{i:06d}/num_agvs"] = np.int64(ep.num_agvs)
            out[f"episode_{i:06d}/num_pickers"] = np.int64(ep.num_pickers)
        dst = os.path.splitext(path)[0] + ".npz"
        np.savez_compressed(dst, **out)
        print(f"{path}: {len(eps)} episodes -> {dst}")


if __name__ == "__main__":
    main()
