"""The path bench.py times, pinned to the float64 oracle at the north star's tolerance.

bench.py steps the FUSED ENGINE (`dcgc_gcmodel_train_step`: forward + loss + backward in one C call) with the
tcgen05 TF32x3 GEMMs.  These tests run exactly that call — same model shape, same synthetic batch (seed of
bench.py::make_pool), B = 4096 — and the Tox21-shaped 12-task classification configuration (BASELINE configs[1]),
against `oracle/graphconv_torch.py` evaluated in float64 and float32 on the host:

  * outputs and loss within 1e-5 (relative to the tensor scale),
  * every gradient tensor, anchored on float64 with no flat slack (helpers.py): rms error
    <= max(1e-5, 3 x the fp32 oracle's), max error <= max(1e-5, 6 x the fp32 oracle's) (the oracle's own error depends
    on the host it runs on, helpers.py),
  * the bf16-GEMM mode: outputs and loss within the north star's 2e-2.  Its GRADIENTS are not within 2e-2 per
    tensor and no kernel can make them: rounding every GEMM operand to bfloat16 perturbs the activations by ~1e-2,
    and the BatchNorm backward passes (differences of large sums) amplify that to 0.1 - 0.19 of a tensor's scale
    (measured, profiles/r4a_parity_probe.md) — the kernels themselves hold 1e-5 against a float64 product of the
    bf16-rounded operands (tests/test_gpu_tc.py::test_bf16_mode_gemms).  Asserted for bf16: the whole gradient
    within 25 % in norm and every tensor with cosine > 0.95 to float64.

Structure follows deepchem/models/tests/test_graphconv_torchmodel.py:15-95 (build the model, load known weights,
one forward, compare every output) extended to the loss and every gradient.
"""
import numpy as np
import pytest
import torch

from helpers import assert_fp64_anchored, fp64_anchored_errors, oracle_batch, oracle_fp32_fp64, rel_err
from oracle import graphconv_torch as O

pytestmark = pytest.mark.gpu

CASES = {
    # the bench configuration (bench.py: LAYERS, DENSE, make_pool seeds of rank 0 / batch 0)
    "bench": dict(B=4096, shape="zinc", mol_seed=0, label_seed=0, layers=[128, 128, 128], dense=128, n_tasks=1,
                  mode="regression", missing=0.0),
    # BASELINE configs[1]: Tox21-shaped, 12 tasks x 2 classes, missing labels as zero weights, reference widths
    "tox21": dict(B=512, shape="tox21", mol_seed=21, label_seed=5, layers=[64, 64], dense=128, n_tasks=12,
                  mode="classification", missing=0.25),
}
# small cases (not asserted here; scripts/parity_probe.py): the shapes of smoke() and of tests/test_gpu_engine.py
CASES["small"] = dict(B=70, shape="stress", mol_seed=11, label_seed=2, layers=[64, 64], dense=128, n_tasks=3,
                      mode="regression", missing=0.25)
CASES["small_zinc"] = dict(B=70, shape="zinc", mol_seed=11, label_seed=2, layers=[128, 128, 128], dense=128, n_tasks=3,
                           mode="classification", missing=0.25)
FLOOR = {"tf32x3": 1e-5, "fp32": 1e-5, "bf16": 2e-2}
_ORACLE = {}      # case -> oracle results (the same parameters and batch serve every GEMM mode)


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def run_case(case, gemm_mode):
    """-> dict with the engine's outputs / loss / gradients and the oracle's fp32 / fp64 ones."""
    from deepchem_b200.data import PackedDataset
    from deepchem_b200.graphconvmodel import GraphConvModel
    from deepchem_b200.synthetic import make_labels, make_molecules
    c = CASES[case]
    B, mode = c["B"], c["mode"]
    pm = make_molecules(B, seed=c["mol_seed"], shape=c["shape"])
    if case == "bench":
        pm = pm.pin_memory()      # as bench.py::make_pool: compact shard -> the engine's input_exact path is the one pinned
    y, w = make_labels(B, c["n_tasks"], mode, seed=c["label_seed"], missing=c["missing"])
    torch.manual_seed(0)
    om = O.OracleGraphConvModel(c["n_tasks"], c["layers"], c["dense"], mode=mode, batch_size=B)
    with torch.no_grad():                      # biases / BatchNorm affine off their zero / one initial values
        for p in om.parameters():
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.1)
    m = GraphConvModel(c["n_tasks"], graph_conv_layers=c["layers"], dense_layer_size=c["dense"], mode=mode,
                       batch_size=B, gemm_mode=gemm_mode)
    assert m._engine is not None, "the fused engine must take this configuration"
    m.model.load_state_dict(om.state_dict())
    batch = next(m.default_generator(PackedDataset(pm, y, w), deterministic=True))
    inputs, labels, weights = m._prepare_batch(batch)
    eng = m._engine
    out = torch.empty(B, eng.cfg.n_out, device=m.device)
    loss = eng.train_step(inputs[1]._dcgc_topology, inputs[0], labels[0].contiguous(), weights[0].contiguous(), B,
                          out=out)
    torch.cuda.synchronize()
    if case not in _ORACLE:
        _, mm = oracle_batch(pm.to_list())
        _ORACLE[case] = oracle_fp32_fp64(om, mode, mm, B, batch[1][0], w)
    res = _ORACLE[case]
    k = 1 if mode == "classification" else 0          # logits / regression output
    return dict(model=m, out=out.cpu(), loss=float(loss), res=res, out_idx=k,
                grads={n: p.grad.detach().cpu() for n, p in m.model.named_parameters()})


@pytest.mark.parametrize("case,gemm_mode", [("bench", "tf32x3"), ("tox21", "tf32x3"), ("bench", "fp32"),
                                            ("tox21", "fp32"), ("bench", "bf16"), ("tox21", "bf16")])
def test_engine_step_against_float64_oracle(case, gemm_mode):
    _cuda()
    r = run_case(case, gemm_mode)
    floor = FLOOR[gemm_mode]
    o32, l32, g32 = r["res"][torch.float32]
    o64, l64, g64 = r["res"][torch.float64]
    k = r["out_idx"]
    ref_out = o64[k].numpy()
    e_out = rel_err(r["out"].numpy().reshape(ref_out.shape), ref_out)
    e_loss = abs(r["loss"] - l64) / max(abs(l64), 1e-30)
    print("%s/%s: out %.2e (fp32 oracle %.2e)  loss %.2e (fp32 oracle %.2e)" % (
        case, gemm_mode, e_out, rel_err(o32[k].numpy(), ref_out), e_loss, abs(l32 - l64) / max(abs(l64), 1e-30)))
    assert e_out <= floor and e_loss <= floor
    worst = (0.0, 0.0, None)
    if gemm_mode == "bf16":
        num = sum(float((g.double() - g64[n].double()).pow(2).sum()) for n, g in r["grads"].items())
        den = sum(float(g64[n].double().pow(2).sum()) for n in r["grads"])
        print("%s/bf16: whole-gradient relative L2 error %.3f" % (case, (num / den) ** 0.5))
        assert (num / den) ** 0.5 < 0.25
        for name, g in r["grads"].items():
            ref = g64[name].double()
            if float(ref.norm()) > 0:
                cos = float((g.double() * ref).sum() / (float(g.double().norm()) * float(ref.norm())))
                assert cos > 0.95, (name, cos)
        return
    for name, g in r["grads"].items():
        e, base = assert_fp64_anchored(name, g, g32[name], g64[name], floor=floor)
        if e > worst[0]:
            worst = (e, base, name)
    print("%s/%s: worst gradient %s: %.2e of scale (fp32 oracle %.2e)" % (case, gemm_mode, worst[2], worst[0], worst[1]))


def test_running_statistics_of_the_engine_step():
    """running_mean / running_var after one engine step in TF32x3 mode against the float64 oracle."""
    import copy
    from deepchem_b200.data import PackedDataset
    from deepchem_b200.graphconvmodel import GraphConvModel
    from deepchem_b200.synthetic import make_labels, make_molecules
    from helpers import torch_args
    _cuda()
    B = 1024
    pm = make_molecules(B, seed=2, shape="zinc")
    y, w = make_labels(B, 1, "regression", seed=2)
    torch.manual_seed(1)
    om = O.OracleGraphConvModel(1, [128, 128], 128, mode="regression", batch_size=B)
    m = GraphConvModel(1, [128, 128], 128, mode="regression", batch_size=B, gemm_mode="tf32x3")
    m.model.load_state_dict(om.state_dict())
    batch = next(m.default_generator(PackedDataset(pm, y, w), deterministic=True))
    inputs, labels, weights = m._prepare_batch(batch)
    m._engine.train_step(inputs[1]._dcgc_topology, inputs[0], labels[0].contiguous(), weights[0].contiguous(), B)
    _, mm = oracle_batch(pm.to_list())
    o64 = copy.deepcopy(om).double()
    o64.train()
    o64(torch_args(mm, B, dtype=torch.float64))
    sd, sd64 = m.model.state_dict(), o64.state_dict()
    for k_ in sd:
        if "running" in k_:
            assert rel_err(sd[k_].cpu().numpy(), sd64[k_].numpy()) < 1e-5, k_
