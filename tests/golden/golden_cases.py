"""Small model configs shared by ``make_golden.py`` (which runs the reference on them)
and by the tests (which replay the stored reference outputs).  Constructor kwargs are
the reference's own (vit.py:104-121, rankvit.py:158-175, residualvit.py:390-415,
adavit.py:229-248, moevit.py:210-224)."""

_BASE = dict(image_size=64, patch_size=8, num_layers=4, num_heads=2, hidden_dim=128, mlp_dim=256, num_classes=10)

CASES = {
    # dh = 32, odd token count (17)
    "vit_d64_h2": dict(
        family="vit", batch=4, weight_seed=11, image_seed=21,
        cfg=dict(image_size=32, patch_size=8, num_layers=3, num_heads=2, hidden_dim=64, mlp_dim=128, num_classes=10)),
    # dh = 64, registers + two class tokens (sum readout, vit.py:242-243)
    "vit_d128_regs": dict(
        family="vit", batch=3, weight_seed=12, image_seed=22,
        cfg=dict(image_size=48, patch_size=16, num_layers=2, num_heads=2, hidden_dim=128, mlp_dim=256,
                 num_classes=7, num_registers=2, num_class_tokens=2)),
    "rankvit_b05": dict(
        family="rankvit", batch=4, weight_seed=13, image_seed=23, budget=0.5,
        cfg=dict(_BASE, rankvit_layers=[1, 3])),
    "rankvit_list": dict(
        family="rankvit", batch=4, weight_seed=13, image_seed=24, budget=[1, 0.5, 1, 0.25],
        cfg=dict(_BASE, rankvit_layers=[1, 3])),
    "residual_learnable_b04": dict(
        family="residualvit", batch=4, weight_seed=14, image_seed=25, budget=0.4,
        cfg=dict(_BASE, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5,
                 add_budget_token="learnable", residual_layers=["attention+mlp"] * 4)),
    "residual_learnable_b08": dict(
        family="residualvit", batch=4, weight_seed=14, image_seed=25, budget=0.8,
        cfg=dict(_BASE, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5,
                 add_budget_token="learnable", residual_layers=["attention+mlp"] * 4)),
    # gate biases calibrated so every layer keeps a mid-range fraction
    "residual_learnable_cal04": dict(
        family="residualvit", batch=4, weight_seed=18, image_seed=29, budget=0.4, calibrate=0.4,
        cfg=dict(_BASE, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5,
                 add_budget_token="learnable", residual_layers=["attention+mlp"] * 4)),
    "residual_fixed_b05": dict(
        family="residualvit", batch=3, weight_seed=15, image_seed=26, budget=0.5,
        cfg=dict(_BASE, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5,
                 add_budget_token=True, residual_layers=["attention+mlp"] * 4)),
    "avit": dict(
        family="adavit", batch=4, weight_seed=16, image_seed=27,
        cfg=dict(_BASE, num_layers=6, eps=0.01, gate_scale=3.0, gate_center=-0.3)),
    "moevit": dict(
        family="moevit", batch=4, weight_seed=17, image_seed=28,
        cfg=dict(_BASE, num_layers=3, mlp_moes=[1, 4, 2])),
    # attention experts (moevit.py:71-102) on layers 0 and 2, expert MLPs on layer 1
    "moevit_attn": dict(
        family="moevit", batch=4, weight_seed=21, image_seed=31,
        cfg=dict(_BASE, num_layers=3, mlp_moes=[1, 2, 1], attn_moes=[2, 1, 3])),
    # early-exit heads after every residual layer (eeresidualvit.py:73-96); outputs = exits + [final logits]
    "eeresidual_learnable_cal04": dict(
        family="eeresidualvit", batch=4, weight_seed=19, image_seed=30, budget=0.4, calibrate=0.4,
        cfg=dict(_BASE, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5,
                 add_budget_token="learnable", residual_layers=["attention+mlp"] * 4)),
    # NoiseBlock spliced in front of a block by add_noise (utils/utils.py:162-191; blocks.py:100-188)
    "vit_noise_snr": dict(
        family="vit", batch=3, weight_seed=22, image_seed=32, noise_seed=77,
        noise=dict(layer=1, noise_type="gaussian", snr=10.0),
        cfg=dict(_BASE, num_layers=3)),
    "vit_noise_token_drop": dict(
        family="vit", batch=3, weight_seed=22, image_seed=32, noise_seed=78,
        noise=dict(layer=2, noise_type="token_drop", prob=0.3),
        cfg=dict(_BASE, num_layers=3)),
    # ---- the ResidualViT configurations that run on the dense masked row layout (Forward.residualvit_dense)
    # skip mode 'attention' (residualvit.py:130-157): no budget token possible, the gate's own threshold
    "residual_skip_attention": dict(
        family="residualvit", batch=3, weight_seed=31, image_seed=41,
        cfg=dict(_BASE, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5,
                 residual_layers=["attention", None, "attention", "none"])),
    # skip mode 'mlp' (:160-194) thresholds on the batch mean of the (fixed-float) budget token; mixed with the other modes
    "residual_skip_mlp_fixed": dict(
        family="residualvit", batch=3, weight_seed=32, image_seed=42, budget=0.5,
        cfg=dict(_BASE, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5,
                 add_budget_token=True, residual_layers=["attention+mlp", "mlp", None, "mlp"])),
    # 'mlp' with add_input (:189-192; only without a budget token) followed by an 'attention' layer
    "residual_skip_mlp_add_input": dict(
        family="residualvit", batch=3, weight_seed=33, image_seed=43,
        cfg=dict(_BASE, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5, add_input=True,
                 residual_layers=["mlp", None, "mlp", "attention"])),
    # gumbel gate in eval = round(sigmoid(logit)) (blocks.py:55-57)
    "residual_gumbel_modes": dict(
        family="residualvit", batch=3, weight_seed=34, image_seed=44,
        cfg=dict(_BASE, gate_type="gumbel", residual_layers=["attention", "mlp", None, "mlp"])),
    # two class tokens + a register: the further class token and the register are gated like image tokens, the head sums both
    # class tokens (:609-611)
    "residual_two_cls_cal05": dict(
        family="residualvit", batch=4, weight_seed=35, image_seed=45, budget=0.5, calibrate=0.5,
        cfg=dict(_BASE, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5, num_class_tokens=2,
                 num_registers=1, add_budget_token="learnable", residual_layers=["attention+mlp"] * 4)),
    # NoiseBlock inside a ResidualViT encoder (utils/utils.py:162-191): dropped rows stop being identical
    "residual_noise_snr": dict(
        family="residualvit", batch=3, weight_seed=18, image_seed=46, budget=0.4, calibrate=0.4, noise_seed=79,
        noise=dict(layer=2, noise_type="gaussian", snr=8.0),
        cfg=dict(_BASE, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5,
                 add_budget_token="learnable", residual_layers=["attention+mlp"] * 4)),
    "residual_noise_token_drop": dict(
        family="residualvit", batch=3, weight_seed=18, image_seed=47, budget=0.4, calibrate=0.4, noise_seed=80,
        noise=dict(layer=1, noise_type="token_drop", prob=0.25),
        cfg=dict(_BASE, gate_type="sigmoid", gate_temp=1.0, gate_bias=0.0, gate_threshold=0.5,
                 add_budget_token="learnable", residual_layers=["attention+mlp"] * 4)),
}


def oracle_noise(case, device="cpu"):
    """The random draw the NoiseBlock of ``case`` makes, regenerated from ``noise_seed`` on ``device``'s torch generator,
    in the form ``oracle.peekvit_oracle.noise_block`` takes.  Call order matters: this must be the first consumer of the
    generator after seeding, exactly like the forward it is compared with."""
    import torch
    nz, cfg = case["noise"], case["cfg"]
    n_tok = (cfg["image_size"] // cfg["patch_size"]) ** 2 + cfg.get("num_class_tokens", 1) + cfg.get("num_registers", 0)
    if cfg.get("add_budget_token"):
        n_tok += 1                                     # ResidualViT: the budget token is part of the sequence the block sees
    torch.manual_seed(case["noise_seed"])
    if nz["noise_type"] == "gaussian":
        return dict(layer=nz["layer"], snr_db=nz["snr"],
                    noise=torch.randn(case["batch"], n_tok, cfg["hidden_dim"], device=device).cpu())
    return dict(layer=nz["layer"], prob=nz["prob"], perm=torch.randperm(n_tok))


def build_case(case):
    """(state_dict, images) of a case, regenerated from its seeds."""
    from oracle import weights as ow
    sd = ow.make_state_dict(case["family"], case["cfg"], seed=case["weight_seed"])
    if case.get("calibrate") is not None:
        sd = ow.calibrate_residual_gates(sd, case["cfg"], case["calibrate"])
    images = ow.synthetic_images(case["batch"], case["cfg"]["image_size"], seed=case["image_seed"])
    return sd, images
