from .compile import compile_brevitas_qat_model, compile_torch_model  # noqa: F401
