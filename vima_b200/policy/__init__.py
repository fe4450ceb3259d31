from .vima_policy import VIMAPolicy
