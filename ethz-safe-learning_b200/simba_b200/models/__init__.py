from .model import BaseModel
from .mlp_ensemble import MlpEnsemble
from .transition_model import TransitionModel
