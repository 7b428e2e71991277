"""Generate tests/golden/value_nets.npz from the reference's shipped value networks, BY RUNNING THE REFERENCE'S OWN
`mlp` class (model.py:14-51) on them.

Run in the build container only (needs /root/reference):  python tests/golden/make_value_nets.py
For each game_theoretic_NN/models/V_GT_sc{k}.pt (k = 1..8) the checkpoint is loaded into the reference's module
exactly as mpc.py:119-125 does (hidden_size / num_layers from configs/sc{k}_config.yaml), evaluated in fp64 on 64
seeded inputs, and the weights, inputs and outputs are stored.  The dataset the reference whitens its features with
(mpc.py:108-118) is not shipped, so the golden inputs are the network's raw inputs.
"""
import os
import sys
import types

import numpy as np
import torch
import yaml

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
REF = os.environ.get("IGT_REFERENCE_ROOT", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.modules.setdefault("casadi", types.ModuleType("casadi"))          # model.py imports it for get_casadi_mlp only
    sys.path.insert(0, REF)
    import model as ref_model                                             # the reference's model.py
    rng = np.random.default_rng(2026)
    out = {}
    for k in range(1, 9):
        cfg = yaml.safe_load(open(os.path.join(REF, "game_theoretic_NN", "configs", "sc%d_config.yaml" % k)))
        net = ref_model.mlp(input_layer_size=cfg["input_size"], output_layer_size=1,
                            hidden_layer_sizes=[cfg["hidden_size"]] * cfg["num_layers"], activation='tanh', batch_norm=False)
        sd = torch.load(os.path.join(REF, "game_theoretic_NN", "models", "V_GT_sc%d.pt" % k), map_location="cpu")
        net.load_state_dict(sd)
        net.eval()
        lin = [m for m in net.mlp if isinstance(m, torch.nn.Linear)]
        out["sc%d_n_layers" % k] = np.array(len(lin))
        for i, m in enumerate(lin):
            out["sc%d_W%d" % (k, i)] = m.weight.detach().numpy().astype(np.float64)
            out["sc%d_b%d" % (k, i)] = m.bias.detach().numpy().astype(np.float64)
        x = rng.normal(size=(64, cfg["input_size"])) * np.array([1.0, 1.0, 1.0, 1.0, 1.0, 1.0])
        with torch.no_grad():
            y = net(torch.from_numpy(x)).numpy()[:, 0]
        out["sc%d_x" % k], out["sc%d_y" % k] = x, y
    np.savez_compressed(os.path.join(OUT, "value_nets.npz"), **out)
    print("wrote value_nets.npz:", {k: int(out["sc%d_n_layers" % k]) for k in range(1, 9)})


if __name__ == "__main__":
    main()
