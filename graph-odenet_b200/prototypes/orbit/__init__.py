from .model import MLP, IN, IN_ODEfunc, IN_ODE  # noqa: F401
