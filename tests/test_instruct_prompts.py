"""Retrieval query builders and the instruction prompts under them (host code: strings and index lists) against the
output of the UNMODIFIED reference functions (tests/golden/make_golden.py::golden_retrieval_inputs):
procyon/data/inference_utils.py:663-925, procyon/data/instruct_tune/instruct_constructor.py:111-366."""
import glob
import hashlib
import json
import os
import types

import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "retrieval_inputs.pt")
REF_TASKS = "/root/reference/procyon/data/instruct_tune/tasks"


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


def _same(a, b, path="root"):
    if isinstance(a, dict):
        assert isinstance(b, dict) and a.keys() == b.keys(), path
        for k in a:
            _same(a[k], b[k], f"{path}.{k}")
    elif isinstance(a, (list, tuple)):
        assert isinstance(b, (list, tuple)) and len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, f"{path}[{i}]")
    elif isinstance(a, torch.Tensor):
        assert isinstance(b, torch.Tensor) and a.dtype == b.dtype and torch.equal(a, b), path
    else:
        assert a == b, f"{path}: {a!r} != {b!r}"


def test_prompts_match_reference_on_synthetic_tasks(gold):
    from procyon_b200.data.instruct_tune.instruct_constructor import get_prompt, get_prompt_open_def

    n = 0
    for (name, n_ex, kind, variant), want in gold["prompts"].items():
        t = gold["tasks"][name]
        ppi = t["DATASET_IDENTIFIER"] == "protein"
        if variant == "fixed":
            got = get_prompt(t, num_examples=n_ex, is_special_definition=False, is_ppi=ppi, aaseq_type=kind)
        elif variant == "open":
            got = get_prompt_open_def(t, num_examples=n_ex, is_special_definition=False, is_ppi=ppi, aaseq_type=kind)
        else:
            got = get_prompt(t, num_examples=n_ex, is_special_definition=True, is_ppi=ppi, aaseq_type=kind)
        _same(tuple(want), tuple(got), f"{name}/{n_ex}/{kind}/{variant}")
        n += 1
    assert n == len(gold["prompts"]) and n > 200


@pytest.mark.skipif(not os.path.isdir(REF_TASKS), reason="the reference checkout is not present")
def test_prompts_match_reference_on_its_own_task_files(gold):
    """every (task file, example count, sequence kind) of the reference's 66 task descriptions, by digest"""
    from procyon_b200.data.instruct_tune.instruct_constructor import get_prompt, get_prompt_open_def

    checked = 0
    for path in sorted(glob.glob(os.path.join(REF_TASKS, "*.json"))):
        t = json.load(open(path))
        name = os.path.basename(path)[:-5]
        ppi = name.startswith("protein_") or name.startswith("domain_protein_")
        for n_ex in (0, 1, 2):
            for kind in ("protein", "domain"):
                try:
                    out = (get_prompt(t, num_examples=n_ex, is_ppi=ppi, aaseq_type=kind),
                           get_prompt_open_def(t, num_examples=n_ex, is_ppi=ppi, aaseq_type=kind))
                except Exception:
                    out = "error"
                want = gold["task_file_digests"][f"{name}|{n_ex}|{kind}"]
                assert hashlib.sha256(repr(out).encode()).hexdigest() == want, (name, n_ex, kind)
                checked += 1
    assert checked == len(gold["task_file_digests"])


def test_retrieval_input_builders_match_reference(gold, tmp_path):
    import pandas as pd

    from procyon_b200.data.inference_utils import create_batched_input_retrieval, create_input_retrieval

    home, data = tmp_path / "home", tmp_path / "data"
    tdir = home / "procyon" / "data" / "instruct_tune" / "tasks"
    tdir.mkdir(parents=True)
    for name, t in gold["tasks"].items():
        (tdir / f"{name}.json").write_text(json.dumps(t))
    for ds, cols in gold["tables"].items():
        d = data / "integrated_data" / "v1" / ds
        d.mkdir(parents=True)
        pd.DataFrame(cols).to_pickle(d / f"{ds}_info_filtered_composed.pkl")
    torch.save(gold["drug_mask"], data / "integrated_data" / "v1" / "drugbank" / "drugbank_mask.pt")
    da = types.SimpleNamespace(retrieval_subset_version=1)
    roots = dict(home_dir=str(home), data_dir=str(data))
    for kw, want in zip(gold["calls"], gold["single"]):
        got = create_input_retrieval(data_args=da, **kw, **roots)
        _same(want, got, f"single{kw}")
    got = create_batched_input_retrieval(data_args=da, **gold["batched_kw"], **roots)
    _same(gold["batched"], got, "batched")
    # the environment roots work like the reference's module-level HOME_DIR / DATA_DIR
    os.environ["HOME_DIR"], old = str(home), os.environ.get("HOME_DIR")
    os.environ["DATA_DIR"], old_d = str(data), os.environ.get("DATA_DIR")
    try:
        _same(gold["single"][0], create_input_retrieval(data_args=da, **gold["calls"][0]), "env roots")
    finally:
        for k, v in (("HOME_DIR", old), ("DATA_DIR", old_d)):
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_caption_and_qa_input_builders_match_reference(gold, tmp_path):
    import pandas as pd

    from procyon_b200.data.inference_utils import create_caption_input_simple, create_qa_input_simple

    home, data = tmp_path / "home", tmp_path / "data"
    tdir = home / "procyon" / "data" / "instruct_tune" / "tasks"
    tdir.mkdir(parents=True)
    for name, t in gold["tasks"].items():
        (tdir / f"{name}.json").write_text(json.dumps(t))
    for ds, cols in gold["tables"].items():
        d = data / "integrated_data" / "v1" / ds
        d.mkdir(parents=True)
        pd.DataFrame(cols).to_pickle(d / f"{ds}_info_filtered_composed.pkl")
    da = types.SimpleNamespace(qa_subset_version=1, caption_subset_version=1)
    roots = dict(home_dir=str(home), data_dir=str(data),
                 functional_descriptions=pd.Series([f"function of sequence {i}" for i in range(80)]))
    for kw, want in zip(gold["cap_calls"], gold["caption"]):
        _same(want, create_caption_input_simple(data_args=da, **kw, **roots), f"caption{kw}")
    for kw, want in zip(gold["qa_calls"], gold["qa"]):
        _same(want, create_qa_input_simple(data_args=da, **kw, **roots), f"qa{kw}")
    # upstream, the QA builder cannot take a task_definition (the "{answer}" slot is still open when the definition
    # is formatted in): same failure here
    with pytest.raises(KeyError):
        create_qa_input_simple(data_args=da, **gold["qa_calls"][0], task_definition="x", **roots)


def test_subset_tables_equal_reference_constants(gold):
    """every column table, every version, against the copies taken from the reference's constants module"""
    from procyon_b200.data.inference_utils import CAPTION_SUBSETS, QA_SUBSETS, RETRIEVAL_SUBSETS

    _same(gold["subsets"]["retrieval"], RETRIEVAL_SUBSETS, "RETRIEVAL_SUBSETS")
    _same(gold["subsets"]["qa"], QA_SUBSETS, "QA_SUBSETS")
    _same(gold["subsets"]["caption"], CAPTION_SUBSETS, "CAPTION_SUBSETS")


def test_retrieval_subsets_match_reference_constants():
    """column names are the contract with the reference's data files (procyon/data/constants.py:241-330)"""
    from procyon_b200.data.inference_utils import RETRIEVAL_SUBSETS

    assert set(RETRIEVAL_SUBSETS) == {1, 2, 5}
    assert RETRIEVAL_SUBSETS[1]["disgenet"][:3] == ["description_air", "description_aot", "description_chv"]
    assert len(RETRIEVAL_SUBSETS[1]["disgenet"]) == 18 and RETRIEVAL_SUBSETS[2]["disgenet"] == ["description_all_collapse"]
    assert RETRIEVAL_SUBSETS[5]["omim"] == ["omim_def_curated", "omim_clinical_curated", "omim_molecular_curated",
                                            "omim_title_curated"]
    assert RETRIEVAL_SUBSETS[5]["pfam"] == RETRIEVAL_SUBSETS[1]["pfam"] == ["description_pfam", "description_interpro"]
