from .procyon import (  # noqa: F401
    EvalArgs,
    ProcyonCaptionEval,
    ProcyonQAEval,
    ProcyonRetrievalEval,
    model_zoo,
    move_inputs_to_device,
)
