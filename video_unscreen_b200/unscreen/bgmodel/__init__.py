from .agent import BackgroundAgent  # noqa
