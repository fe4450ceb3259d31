from .vima_policy import VIMAPolicy
from .vima_gato_policy import VIMAGatoPolicy
from .vima_gpt_policy import VIMAGPTPolicy
from .vima_flamingo_policy import VIMAFlamingoPolicy
