"""CPU oracle for the procyon_b200 hot path — TEST INFRASTRUCTURE ONLY.

A plain PyTorch (CPU, fp32) restatement of the reference algorithm for the protein-text fusion forward
path of mims-harvard/ProCyon. Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import this package; the product (`procyon_b200/`) never does.

Parity status: the reference's own test-suite pins nothing on this path (its only test file,
procyon/evaluate/framework/testing.py, covers CPU metric plumbing) and the arithmetic lives in un-vendored
third-party packages (fair-esm 2.0.0, transformers 4.31.0, torch 2.2.0 — pyproject.toml:8-55).  The oracle
is therefore pinned two ways, both committed under tests/golden/ with the generating script
(tests/golden/make_golden.py):
  * host-level functions that DO live in the reference tree (ProteinPooler, batched_split_long_seq,
    reverse_batched_split, create_mlp, InfoNCEInBatch, left_pad_tensors, compute_conflict_matrix,
    mask_before, multi_replace_tokens, _generate_beam_search, _get_nucleus_mask, _prepare_input_embeddings)
    are executed from /root/reference itself (imported with stubs for the absent third-party modules) and
    their outputs stored as golden vectors;
  * the two transformer stacks are cross-checked against the installed HuggingFace `EsmForMaskedLM` /
    `LlamaForCausalLM` (transformers 5.5.0, eager attention, seeded random weights), whose ESM2 and Llama
    math is the same as the pinned fair-esm / transformers-4.31 code on the inputs used.
"""
