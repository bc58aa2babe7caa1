"""Import shadow of the reference's `model` package.

The reference's scripts bind the hot path with `from model.model_tools import load_model, total_parameters_count`
(train_and_evaluate_sp.py:20, train_and_evaluate_wp.py:20, demo/demo.py:13, utils/visualization.py:11) and
`from model.KASportsFormer import KASportsFormer`.  Putting this directory BEFORE the reference checkout on the
module search path

    PYTHONPATH=/path/to/kasportsformer_b200/shadow:/path/to/KASportsFormer python train_and_evaluate_sp.py --config-path ...

swaps the B200 implementation in with the scripts unchanged: same constructor, same YAML keys, same `state_dict`
names (so `load_state_dict(strict=True)` of a released, `module.`-prefixed checkpoint on the `nn.DataParallel` wrap of
train_and_evaluate_sp.py:162-176 works), same `forward(x, return_rep=False)`.  Inference only: `model.train()` +
forward raises.  (tests/test_dropin_cpu.py imports the unmodified script this way.)
"""
