from .eval import (compute_EIG_from_history, compute_ll, compute_rmse, eval_boed,  # noqa: F401
                   eval_EIG_from_history, get_traces)
from .target_mask import create_target_mask, select_targets_by_mask  # noqa: F401
from .misc import calculate_gmm_variance  # noqa: F401
