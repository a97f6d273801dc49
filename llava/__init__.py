"""Import shim so the reference's eval drivers run UNCHANGED against the B200 engine.

The drivers do (reference experiments/eval/sampling/llava_sampling.py:12-16, eval/calibrate/llava_calibrate.py:12-16,
eval/MME/run_llava.py:10-14):

    from llava.constants import IMAGE_TOKEN_INDEX, DEFAULT_IMAGE_TOKEN, ...
    from llava.conversation import conv_templates, SeparatorStyle
    from llava.model.builder import load_pretrained_model
    from llava.utils import disable_torch_init
    from llava.mm_utils import tokenizer_image_token, get_model_name_from_path, KeywordsStoppingCriteria

With this repo's root first on PYTHONPATH, `llava.model.builder.load_pretrained_model` resolves to
llava_align_b200.builder (same signature and return tuple, reference llava/model/builder.py:26,148).  The host-side
helpers stay the reference's own files when LLAVA_ALIGN_REFERENCE_LLAVA points at its `experiments/llava` directory (that
directory is appended to this package's search path: `llava.conversation`, `llava.mm_utils`, `llava.utils`,
`llava.constants` then load from there, untouched; the reference's own `llava/__init__.py`, which imports its HF model
classes, is never executed); otherwise llava_align_b200.hostside serves them.
"""
import os

_ref = os.environ.get("LLAVA_ALIGN_REFERENCE_LLAVA")
if _ref and os.path.isdir(_ref):
    __path__.append(_ref)
else:
    from llava_align_b200 import hostside as _hostside
    _hostside.register_fallback_modules(__name__)

from .model import LlavaLlamaForCausalLM  # noqa: E402,F401  (reference llava/__init__.py:1)
