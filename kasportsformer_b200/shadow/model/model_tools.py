"""model/model_tools.py of the reference -> the B200 drop-in (reference :79-104)."""
from kasportsformer_b200.model import load_model, total_parameters_count  # noqa: F401
