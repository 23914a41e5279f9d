"""Minimal stand-in for sacred (test infrastructure; see ../README.md).

Only what src/vilt/config.py touches: Experiment(name) with the .config / .named_config / .automain
decorators.  `materialize(named, overrides)` evaluates the decorated function BODIES in order
(default config, then named configs, then key=value overrides), the way sacred resolves
`python run.py with <named...> k=v`.
"""
import inspect
import textwrap


class Experiment:
    def __init__(self, name=None, **kwargs):
        self.name = name
        self.default_configs = []
        self.named_configs = {}
        self.main = None

    def config(self, fn):
        self.default_configs.append(fn)
        return fn

    def named_config(self, fn):
        self.named_configs[fn.__name__] = fn
        return fn

    def automain(self, fn):
        self.main = fn
        return fn

    @staticmethod
    def _run_body(fn, cfg):
        src = textwrap.dedent(inspect.getsource(fn))
        lines = src.splitlines()
        start = next(i for i, line in enumerate(lines) if line.lstrip().startswith("def "))
        body = textwrap.dedent("\n".join(lines[start + 1:]))
        env = dict(fn.__globals__)
        env.update(cfg)
        local = {}
        exec(compile(body, inspect.getsourcefile(fn) or "<config>", "exec"), env, local)
        for k, v in local.items():
            if not k.startswith("_"):
                cfg[k] = v

    def materialize(self, named=(), overrides=None):
        cfg = {}
        for fn in self.default_configs:
            self._run_body(fn, cfg)
        for name in named:
            self._run_body(self.named_configs[name], cfg)
        cfg.update(overrides or {})
        return cfg
