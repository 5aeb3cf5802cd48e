from . import seeding, passive_env_checker


class RecordConstructorArgs:
    def __init__(self, *, _disable_deepcopy: bool = False, **kwargs):
        if not hasattr(self, "_saved_kwargs"):
            self._saved_kwargs = kwargs
