"""Plug-in boundary: the Signal / Module / Network protocol the hot-path modules live in.

The reference's modular runtime (pymoto/core_objects.py:71-195 Signal, :456-744 Module, :750-994 Network) is the
plug-in API and is NOT re-implemented as a product: when pyMOTO is installed, the classes of this package derive
from ``pymoto.Module`` and are used inside a ``pymoto.Network`` next to the reference's own modules.  When it
is not importable (the GPU box; this image lacks matplotlib, which pymoto imports at module level), the small
stand-ins below provide the same protocol so that the same user script runs:

  * ``Mod(cfg)(sig_a, sig_b)`` connects the module to its input signals, evaluates it once, creates the output
    ``Signal``s, registers the module in every active ``Network`` and returns the output signal(s); called with
    plain values it behaves as a function (core_objects.py:494-546).
  * ``response()`` recomputes outputs from inputs; ``sensitivity()`` maps output sensitivities to input
    sensitivities through ``_sensitivity`` and accumulates them (``None`` = no contribution) (:625-674);
    ``reset()`` clears sensitivities (:676-690).
  * ``Network`` evaluates modules in insertion order and back-propagates in reverse (:803-829).
"""
import copy
import time

try:  # pragma: no cover - exercised only where pyMOTO is installed
    import pymoto as _pym

    Signal, Module, Network = _pym.Signal, _pym.Module, _pym.Network
    HAVE_PYMOTO = True
except Exception:  # ModuleNotFoundError (pymoto or its matplotlib import)
    HAVE_PYMOTO = False

    def _as_list(v):
        if v is None:
            return []
        if isinstance(v, (list, tuple)):
            return list(v)
        return [v]

    class Signal:
        def __init__(self, tag="", state=None, sensitivity=None, min=None, max=None):
            self.tag, self.state, self.sensitivity, self.min, self.max = tag, state, sensitivity, min, max
            self.keep_alloc = sensitivity is not None

        def add_sensitivity(self, ds):
            if ds is None:
                return self
            if self.sensitivity is None:
                self.sensitivity = copy.deepcopy(ds)
            elif hasattr(self.sensitivity, "add_sensitivity"):
                self.sensitivity.add_sensitivity(ds)
            else:
                self.sensitivity += ds
            return self

        def reset(self, keep_alloc=None):
            if self.sensitivity is None:
                return self
            if self.keep_alloc if keep_alloc is None else keep_alloc:
                self.sensitivity[...] = 0
            else:
                self.sensitivity = None
            return self

        def __repr__(self):
            return f'Signal "{self.tag}"'

    def _is_signal(s):
        return all(hasattr(s, f) for f in ("state", "sensitivity", "add_sensitivity", "reset"))

    class Module:
        sig_in = None
        sig_out = None

        def __init_subclass__(cls, **kwargs):
            super().__init_subclass__(**kwargs)
            fwd = cls.__dict__.get("__call__")
            if fwd is None or getattr(fwd, "_pmb_wrapped", False):
                return

            def connect_and_call(self, *args, _fwd=fwd):
                states = [a.state if _is_signal(a) else a for a in args]
                if len(args) > 0 and not any(_is_signal(a) for a in args):
                    return _fwd(self, *states)  # plain function call
                self.sig_in = list(args)
                out = _as_list(_fwd(self, *states))
                if self.sig_out is None:
                    self.sig_out = [Signal(f"{type(self).__name__}_output{i}") for i in range(len(out))]
                    for n in Network.active:
                        n.append(self)
                for s, v in zip(self.sig_out, out):
                    s.state = v
                return None if not self.sig_out else (self.sig_out[0] if len(self.sig_out) == 1 else tuple(self.sig_out))

            connect_and_call._pmb_wrapped = True
            cls._orig_call = fwd
            cls.__call__ = connect_and_call

        def response(self):
            try:
                out = _as_list(self._orig_call(*[s.state if _is_signal(s) else s for s in self.sig_in]))
                for s, v in zip(self.sig_out, out):
                    s.state = v
                return self
            except Exception as e:
                raise type(e)(f"{e}\n\t| raised in response() of module {type(self).__name__}") from e

        def sensitivity(self):
            try:
                dout = [s.sensitivity if _is_signal(s) else None for s in self.sig_out]
                if len(dout) > 0 and all(d is None for d in dout):
                    return self
                din = _as_list(self._sensitivity(*dout))
                if len(din) != len(self.sig_in):
                    raise TypeError(f"Number of sensitivities calculated ({len(din)}) is unequal to number of input "
                                    f"signals ({len(self.sig_in)})")
                for s, d in zip(self.sig_in, din):
                    if _is_signal(s):
                        s.add_sensitivity(d)
                return self
            except Exception as e:
                raise type(e)(f"{e}\n\t| raised in sensitivity() of module {type(self).__name__}") from e

        def reset(self):
            for s in (self.sig_out or []) + (self.sig_in or []):
                if _is_signal(s):
                    s.reset()
            self._reset()
            return self

        def get_input_states(self, as_list=False):
            st = [s.state if _is_signal(s) else s for s in self.sig_in]
            return st[0] if (len(st) == 1 and not as_list) else st

        def _sensitivity(self, *dout):
            return [None for _ in self.sig_in]

        def _reset(self):
            pass

    class Network:
        active = []

        def __init__(self, *mods, print_timing=False):
            self.mods = []
            for m in mods:  # modules or lists of modules (core_objects.py:775-780)
                self.mods.extend(_as_list(m))
            self.print_timing = print_timing

        def __len__(self):
            return len(self.mods)

        def __iter__(self):
            return iter(self.mods)

        def __getitem__(self, i):
            return self.mods[i]

        def __enter__(self):
            Network.active.append(self)
            return self

        def __exit__(self, *exc):
            Network.active.remove(self)

        def append(self, *mods):
            self.mods.extend(mods)

        def _timed(self, mods, what):
            for m in mods:
                t0 = time.time()
                getattr(m, what)()
                if self.print_timing:
                    print(f"{type(m).__name__}.{what}: {time.time() - t0:.4f} s")

        def response(self):
            self._timed(self.mods, "response")

        def sensitivity(self):
            self._timed(list(reversed(self.mods)), "sensitivity")

        def reset(self):
            for m in reversed(self.mods):
                m.reset()

        @staticmethod
        def _signal_set(sigs):
            return {id(s) for s in _as_list(sigs) if _is_signal(s)}

        def get_input_cone(self, fromsig=None, frommod=None):
            """Modules that depend on ``fromsig`` (or follow ``frommod``), in evaluation order (core_objects.py:913-928)."""
            touched, frommod = self._signal_set(fromsig), {id(m) for m in _as_list(frommod)}
            if not touched and not frommod:
                return self
            cone = Network(print_timing=self.print_timing)
            for m in self.mods:
                if id(m) in frommod or (self._signal_set(m.sig_in) & touched):
                    touched |= self._signal_set(m.sig_out)
                    cone.append(m)
            return cone

        def get_output_cone(self, tosig=None, tomod=None):
            """Modules ``tosig`` (or ``tomod``) depend on, in evaluation order (core_objects.py:930-945)."""
            dependent, tomod = self._signal_set(tosig), {id(m) for m in _as_list(tomod)}
            if not dependent and not tomod:
                return self
            cone = []
            for m in reversed(self.mods):
                if id(m) in tomod or (self._signal_set(m.sig_out) & dependent):
                    dependent |= self._signal_set(m.sig_in)
                    cone.append(m)
            return Network(list(reversed(cone)), print_timing=self.print_timing)
