"""Calibrated gate parameters for the parity tests at BASELINE shapes (configs C and E).

    python tests/golden/make_calibration.py          # -> tests/golden/calibration.json (authoring container, CPU, minutes)

With random-init weights a ResidualViT gate keeps either every token or none, and the A-ViT halting gate of
``configs/model/avit_s_16_224.yaml`` (scale 10, centre 30: tuned for DeiT-S weights) never halts.  The
tests therefore use gates calibrated with the CPU oracle on the seeded weights of ``oracle.weights``:

* config C (``residualdeit_s_16_224.yaml`` kwargs): per budget in {0.2, 0.4, 0.8, 1.0} the twelve
  ``residual_gate.projection.bias`` values that make a probe batch keep ~budget of its image tokens (1.0 -> ~0.97);
* config E (``avit_s_16_224.yaml`` kwargs): the ``gate_center`` at which the probe batch's tokens run ~7 of 12 layers.

The numbers are stored (not the weights: those are regenerated from the seed) so the GPU tests do not spend their time
bisecting on the CPU.
"""
from __future__ import annotations

import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import weights as ow  # noqa: E402
from oracle import peekvit_oracle as po  # noqa: E402

VITS = dict(image_size=224, patch_size=16, num_layers=12, num_heads=6, hidden_dim=384, mlp_dim=1536, num_classes=1000)
CFG_C = dict(VITS, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5, add_budget_token="learnable",
             add_input=False, residual_layers=["attention+mlp"] * 12)
CFG_E_AVIT = dict(VITS, eps=0.01, gate_scale=10, gate_center=30)
SEED = 4321
GATE_STD = 4.0       # wide token scores: the keep / drop decision varies token by token instead of sample by sample
BT_GATE_SCALE = 0.05  # ... and a budget-token gate that does not saturate: its threshold stays near sigmoid(bias) for every sample
                      # (the budget row grows layer after layer; at the constructor's scale its threshold is 0 or 1 per sample)


def config_c_state_dict(seed=SEED, gate_std=GATE_STD, bt_gate_scale=BT_GATE_SCALE):
    sd = ow.make_state_dict("residualvit", CFG_C, seed=seed, gate_std=gate_std)
    for i in range(CFG_C["num_layers"]):
        sd[f"encoder.layers.{i}.budget_token_gate.weight"] = sd[f"encoder.layers.{i}.budget_token_gate.weight"] * bt_gate_scale
    return sd


def main():
    out = {"seed": SEED, "gate_std": GATE_STD, "bt_gate_scale": BT_GATE_SCALE, "config_C": {}, "config_E_avit": {}}
    sd0 = config_c_state_dict()
    probe = ow.synthetic_images(8, 224, seed=99)
    for budget in (0.2, 0.4, 0.8, 1.0):
        sd = ow.calibrate_residual_gates(sd0, CFG_C, min(budget, 0.97), images=probe)
        out["config_C"][str(budget)] = [float(sd[f"encoder.layers.{i}.residual_gate.projection.bias"][0]) for i in range(12)]
        print("C", budget, out["config_C"][str(budget)], flush=True)
    sd = ow.make_state_dict("adavit", CFG_E_AVIT, seed=SEED)
    lo, hi = -40.0, 40.0
    with torch.no_grad():
        for _ in range(16):
            mid = 0.5 * (lo + hi)
            _, aux = po.avit_forward(sd, dict(CFG_E_AVIT, gate_center=mid), probe)
            if float(aux["counter_token"].mean()) > 7.0:
                hi = mid
            else:
                lo = mid
    out["config_E_avit"] = {"gate_scale": 10, "gate_center": 0.5 * (lo + hi), "target_mean_layers": 7.0}
    print("E", out["config_E_avit"], flush=True)
    with open(os.path.join(HERE, "calibration.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
