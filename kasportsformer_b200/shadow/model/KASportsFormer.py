"""model/KASportsFormer.py of the reference -> the B200 drop-in (reference :290-347)."""
from kasportsformer_b200.model import KASportsFormer  # noqa: F401
