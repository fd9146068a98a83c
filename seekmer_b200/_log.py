"""Brace-style logger with the call surface the reference uses from `logbook`
(`Logger.debug/info/warn/warning`, `StderrHandler(...).applicationbound()`), on stdlib logging."""
import contextlib
import logging
import sys


class Logger:
    def __init__(self, name):
        self._log = logging.getLogger(name)

    def _emit(self, level, msg, args, kwargs):
        if self._log.isEnabledFor(level):
            try:
                msg = str(msg).format(*args, **kwargs)
            except (IndexError, KeyError, ValueError):
                pass
            self._log.log(level, msg)

    def debug(self, msg, *a, **k):
        self._emit(logging.DEBUG, msg, a, k)

    def info(self, msg, *a, **k):
        self._emit(logging.INFO, msg, a, k)

    def warn(self, msg, *a, **k):
        self._emit(logging.WARNING, msg, a, k)

    warning = warn

    def error(self, msg, *a, **k):
        self._emit(logging.ERROR, msg, a, k)


class StderrHandler:
    def __init__(self, level='INFO'):
        self.level = getattr(logging, str(level).upper(), logging.INFO)
        self.format_string = None

    @contextlib.contextmanager
    def applicationbound(self):
        handler = logging.StreamHandler(sys.stderr)
        handler.setFormatter(logging.Formatter('%(levelname)-5s %(asctime)s %(name)s: %(message)s',
                                               '%Y-%m-%d %H:%M:%S'))
        root = logging.getLogger('seekmer_b200')
        old = root.level
        root.addHandler(handler)
        root.setLevel(self.level)
        try:
            yield self
        finally:
            root.removeHandler(handler)
            root.setLevel(old)


@contextlib.contextmanager
def nvtx_range(name):
    """An NVTX range around a stage of `infer.run` (visible in Nsight Systems / ncu --nvtx); a
    no-op when torch has not been imported by the caller's process yet or NVTX is unavailable."""
    nvtx = None
    torch = sys.modules.get('torch')
    if torch is not None:
        try:
            nvtx = torch.cuda.nvtx
            nvtx.range_push(name)
        except Exception:
            nvtx = None
    try:
        yield
    finally:
        if nvtx is not None:
            try:
                nvtx.range_pop()
            except Exception:
                pass
