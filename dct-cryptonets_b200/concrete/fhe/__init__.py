"""concrete.fhe.Configuration as used at reference homomorphic_eval.py:266-273 (a plain options object)."""


class Configuration:
    def __init__(self, **options):
        # the reference passes show_progress / progress_tag / progress_title; unknown options are accepted and kept
        self.show_progress = options.pop("show_progress", False)
        self.progress_tag = options.pop("progress_tag", False)
        self.progress_title = options.pop("progress_title", "")
        self.options = options

    def __repr__(self):
        return f"Configuration(show_progress={self.show_progress}, progress_tag={self.progress_tag}, progress_title={self.progress_title!r})"


__all__ = ["Configuration"]
