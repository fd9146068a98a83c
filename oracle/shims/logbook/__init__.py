"""Minimal stand-in for the `logbook` package (absent from this image).

TEST INFRASTRUCTURE ONLY: lets the unmodified reference modules import so the
compiled reference (`oracle/_ref`) can run as the parity oracle / CPU baseline.
Implements just the surface the reference touches.
"""
import contextlib
import sys


class Logger:
    def __init__(self, name=None):
        self.name = name

    def _emit(self, level, msg, *args, **kwargs):
        if not _STATE['enabled']:
            return
        try:
            text = msg.format(*args, **kwargs)
        except Exception:  # pragma: no cover
            text = str(msg)
        print('{:<5} {}: {}'.format(level, self.name, text), file=sys.stderr)

    def debug(self, msg, *a, **k):
        if _STATE['debug']:
            self._emit('DEBUG', msg, *a, **k)

    def info(self, msg, *a, **k):
        self._emit('INFO', msg, *a, **k)

    def warn(self, msg, *a, **k):
        self._emit('WARN', msg, *a, **k)

    warning = warn

    def error(self, msg, *a, **k):
        self._emit('ERROR', msg, *a, **k)


_STATE = {'enabled': False, 'debug': False}


class StderrHandler:
    def __init__(self, level='INFO', **_):
        self.level = level
        self.format_string = None

    @contextlib.contextmanager
    def applicationbound(self):
        old = dict(_STATE)
        _STATE['enabled'] = True
        _STATE['debug'] = self.level == 'DEBUG'
        try:
            yield self
        finally:
            _STATE.update(old)
