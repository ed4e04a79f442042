"""Checkpoint layout compatibility: `model_args.pt` / `data_args.pt` pickled by the REAL reference
dataclasses (tests/golden/ckpt_args, written by make_golden.py from /root/reference) load without the reference
package, stale DATA_DIR prefixes are rewritten like `update_model_args_data_dir` does
(procyon/training/training_args_IT.py:1787-1801), and a whole checkpoint directory round-trips through
`save_pretrained` / `from_pretrained` to bit-identical generations on the GPU."""
import os
import shutil
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CKPT = os.path.join(HERE, "golden", "ckpt_args")


def _load_args():
    from procyon_b200 import compat

    compat.install()
    ma = torch.load(os.path.join(CKPT, "model_args.pt"), weights_only=False)
    da = torch.load(os.path.join(CKPT, "data_args.pt"), weights_only=False)
    exp = torch.load(os.path.join(CKPT, "expected.pt"), weights_only=False)
    return ma, da, exp


def test_reference_pickles_load_into_the_mirror_classes():
    from procyon_b200.training import training_args_IT as T

    ma, da, exp = _load_args()
    assert type(ma) is T.ModelArgs and type(da) is T.DataArgs
    assert "procyon.training.training_args_IT" in sys.modules  # the alias the pickle resolved through
    assert len(vars(ma)) == exp["n_model_fields"] == 93
    for k, v in exp["model_fields"].items():  # every scalar field of the reference instance survives as is
        assert getattr(ma, k) == v, k
    for k, v in exp["data_fields"].items():
        assert getattr(da, k) == v, k
    # the settings UnifiedProCyon reads are llama3-full.yml's
    assert ma.use_aaseq_embeddings and ma.use_protein_struct and ma.use_drug_embeddings and ma.contrastive_global
    assert ma.num_layers_token_projector == 3 and ma.hidden_size_lm_projector == 2560 and ma.ret_token_access == "last"


def test_stale_data_dir_paths_are_rewritten(monkeypatch, tmp_path):
    from procyon_b200.training.training_args_IT import update_data_args_data_dir, update_model_args_data_dir

    ma, da, exp = _load_args()
    old = exp["old_data_dir"]
    stale = [k for k, v in vars(ma).items() if k.endswith("path") and isinstance(v, str) and v.startswith(old)]
    assert len(stale) == exp["n_path_fields"] >= 10
    monkeypatch.setenv("DATA_DIR", str(tmp_path))
    before = dict(vars(ma))
    update_model_args_data_dir(ma, prev_data_dir=da.data_dir)
    for k in stale:
        assert getattr(ma, k) == os.path.join(str(tmp_path), before[k][len(old):].lstrip("/")), k
    for k, v in before.items():
        if k not in stale:
            assert getattr(ma, k) == v, k
    update_data_args_data_dir(da)
    assert da.data_dir == str(tmp_path)
    # same DATA_DIR: nothing to do; wrong type: the reference raises too
    update_model_args_data_dir(ma, prev_data_dir=str(tmp_path))
    with pytest.raises(ValueError):
        update_model_args_data_dir(object(), prev_data_dir=old)


def test_from_pretrained_config_only_and_checkpoint_configs(tmp_path):
    from procyon_b200.model.model_unified import UnifiedProCyon

    for f in ("model_args.pt", "data_args.pt"):
        shutil.copy(os.path.join(CKPT, f), tmp_path / f)
    torch.save({"anything": 1}, tmp_path / "training_args.pt")
    model, cfg = UnifiedProCyon.from_pretrained(checkpoint_dir=str(tmp_path), config_only=True)
    assert model is None and cfg.text_encoder_fname == "llama-3-8b"
    da, ma, ta = UnifiedProCyon.get_checkpoint_configs(str(tmp_path))
    assert ma.max_text_len == 2048 and ta == {"anything": 1}


@pytest.mark.gpu
def test_checkpoint_directory_roundtrip_generates_identically(cuda_device, tmp_path, monkeypatch):
    """save_pretrained -> (txllm_model_ckpt.pt + model_args.pt [a real reference ModelArgs pickle, re-saved] +
    data_args.pt) -> from_pretrained on a machine with another DATA_DIR -> same beams, same scores."""
    from oracle.esm2 import random_protein_tokens
    from procyon_b200.data.simple_tokenizer import SimpleTokenizer
    from procyon_b200.model.model_unified import UnifiedProCyon
    from procyon_b200.model.pmc_llama import LlamaConfig

    ma, da, exp = _load_args()
    # shrink what the real Full config would need from disk (embedding tables, GearNet / drug tables) to a live tiny
    # encoder; everything else stays the reference instance's
    ma.use_aaseq_embeddings = False
    ma.use_protein_struct = False
    ma.use_drug_embeddings = False
    ma.protein_encoder_num_params = "custom"
    ma.max_text_len = 64
    for k in ("hidden_size_token_projector", "hidden_size_shared_projector", "hidden_size_lm_projector"):
        setattr(ma, k, 96)
    lc = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                     num_key_value_heads=1, vocab_size=499, max_position_embeddings=256)
    kw = dict(tokenizer=SimpleTokenizer(base_vocab=499), llama_config=lc, esm_custom_config=(2, 64, 4))
    torch.manual_seed(0)
    m = UnifiedProCyon(ma, **kw)
    for n, p in m.named_parameters():
        if p.dim() > 1:
            torch.nn.init.normal_(p, std=min(0.5, p.shape[1] ** -0.5))
    m = m.bfloat16().eval().cuda()
    m.save_pretrained(str(tmp_path))
    torch.save(da, tmp_path / "data_args.pt")
    assert sorted(os.listdir(tmp_path)) == ["data_args.pt", "model_args.pt", "txllm_model_ckpt.pt"]
    toks = random_protein_tokens(2, 0, seed=8, lengths=[30, 12])
    inputs = {
        "data": {"seq": toks, "seq_idx": torch.tensor([1, 2]), "text": ["binds atp"], "text_idx": [4], "drug": None},
        "input": {"seq": [[1]], "text": [[0]], "drug": None},
        "target": {"seq": None, "text": None, "drug": None},
        "instructions": ["Protein : <|protein|> Context : [EXT] Describe the function . [ANSWER]"],
        "reference_indices": {"input": {"seq": [[5]]}, "target": {"text": [0]}},
    }
    a = m.generate(inputs, max_len=6, method="beam", beam_size=4, beam_group_size=2)
    monkeypatch.setenv("DATA_DIR", str(tmp_path / "another_machine"))
    kw["tokenizer"] = SimpleTokenizer(base_vocab=499)
    m2, cfg2 = UnifiedProCyon.from_pretrained(checkpoint_dir=str(tmp_path), strict_load=True, **kw)
    assert cfg2.protein_seq_embeddings_path.startswith(str(tmp_path / "another_machine"))  # stale prefix rewritten
    m2 = m2.bfloat16().eval().cuda()
    sd1, sd2 = m.state_dict(), m2.state_dict()
    assert sd1.keys() == sd2.keys()
    for k in sd1:
        assert torch.equal(sd1[k], sd2[k]), k
    b = m2.generate(inputs, max_len=6, method="beam", beam_size=4, beam_group_size=2)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and a[3] == b[3]
    assert not b[2].is_cuda  # the logits history is returned on the host, like the reference's (:773-781)
