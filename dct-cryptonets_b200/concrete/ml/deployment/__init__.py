from .fhe_client_server import FHEModelClient, FHEModelDev, FHEModelServer  # noqa: F401
