"""Instruction prompts of the three task families (qa / retrieval / caption) from a task description file.

Mirror of the reference's prompt builder (procyon/data/instruct_tune/instruct_constructor.py:111-330 `get_prompt`,
`get_prompt_open_def`) for the joined-example form (`sample_examples=False`), which is the one the inference input
builders call (procyon/data/inference_utils.py:724-741).  The prompts are assembled from one table of line templates per
(task family, PPI?) instead of the reference's nested f-strings; tests/test_instruct_prompts.py pins every string to the
output of the unmodified reference functions on the reference's own task files.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

_NOUN = {"protein": "Protein", "domain": "Domain", "peptide": "Peptide"}


def aaseq_type_to_prompt(aaseq_type) -> str:
    key = aaseq_type.lower() if isinstance(aaseq_type, str) else aaseq_type
    return _NOUN.get(key, "Amino acid sequence")


# lines of ONE in-context example and of the final instance, per (family, is_ppi); {A} = sequence noun, {out} = yes / no
_EXAMPLE = {
    ("qa", True): "{A} 1: <|protein|>\n{A} 2: <|protein|>\nOutput: [ANSWER] {out}",
    ("qa", False): "Description: [EXT]\n{A}: <|protein|>\n[CONTEXT]Output: [ANSWER] {out}",
    ("retrieval", True): "{A} 1: <|protein|>\n{A} 2: <|protein|>",
    ("retrieval", False): "[CONTEXT]Description: [EXT]\n{A}: <|protein|>",
    ("caption", False): "[CONTEXT]{A}: <|protein|>\nOutput: [ANSWER] [EXT]",
}
_INSTANCE = {
    ("qa", True): "{A} 1: <|protein|>\n{A} 2: <|protein|>\nOutput: [ANSWER] {{answer}}",
    ("qa", False): "Description: [EXT]\n{A}: <|protein|>\n[CONTEXT]Output: [ANSWER] {{answer}}",
    ("retrieval", True): "{A} 1: <|protein|> \n{A} 2: [PROT]",
    ("retrieval", False): "[CONTEXT]Description: [EXT]\n{A}: [PROT]",
    ("caption", False): "[CONTEXT]{A}: <|protein|>\nOutput: [ANSWER] [EXT]",
}


def _examples(examples, positive: bool, family: str, n: Optional[int], is_ppi: bool, noun: str):
    """joined example block + the text / sequence ids it refers to (in order of appearance)"""
    take = examples if n is None else examples[:n]
    head = "Positive example" if positive else "Negative example"
    body = _EXAMPLE[(family, is_ppi)].format(A=noun, out="yes" if positive else "no")
    block = "\n".join(f"{head} {i + 1}:\n{body}" for i in range(len(take)))
    if is_ppi:
        return block, [], [s for e in take for s in (e["aaseq_1"], e["aaseq_2"])]
    return block, [e["text"] for e in take], [e["aaseq"] for e in take]


def _fill_definition(task) -> str:
    out = task["Definition"]
    for key in ("Relationship Summary", "Biological Summary", "Task-Specific Relationship"):
        out = out.replace("{" + key + "}", task[key])
    return out


def _compose(task, definition: str, num_examples, is_ppi: bool, sample_examples: bool, aaseq_type):
    if sample_examples:
        raise NotImplementedError("per-sample example lists (training collator path) are outside the hot path")
    family = task["CATEGORY"]
    if family == "caption":
        assert not is_ppi, "Cannot use PPI with caption task"
    noun = aaseq_type_to_prompt(aaseq_type)
    pos, text_ids, seq_ids = _examples(task["Positive Examples"], True, family, num_examples, is_ppi, noun)
    neg = None
    parts = [f"Definition: {definition}", pos]
    if family == "qa":  # only yes/no questions show negative examples
        neg, nt, ns = _examples(task["Negative Examples"], False, family, num_examples, is_ppi, noun)
        text_ids, seq_ids = text_ids + nt, seq_ids + ns
        parts.append(neg)
    parts.append("Now, complete the following instance:")
    parts.append(_INSTANCE[(family, is_ppi)].format(A=noun))
    return "\n".join(parts), pos, neg, text_ids, seq_ids


def get_prompt(task, num_examples=None, is_special_definition=False, is_ppi=False, sample_examples=False,
               aaseq_type=None) -> Tuple[str, str, Optional[str], List[int], List[int]]:
    """-> (prompt, positive example block, negative example block | None, example text ids, example sequence ids)"""
    definition = task["Definition"] if is_special_definition else _fill_definition(task)
    return _compose(task, definition, num_examples, is_ppi, sample_examples, aaseq_type)


def get_prompt_open_def(task, num_examples=None, is_special_definition=False, is_ppi=False, sample_examples=False,
                        aaseq_type=None):
    """Same prompt with a `{definition}` placeholder; also returns the task's own definition (second element)."""
    definition_true = task["Definition"] if is_special_definition else _fill_definition(task)
    prompt, pos, neg, text_ids, seq_ids = _compose(task, "{definition}", num_examples, is_ppi, sample_examples, aaseq_type)
    return prompt, definition_true, pos, neg, text_ids, seq_ids
