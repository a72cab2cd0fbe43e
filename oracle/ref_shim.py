"""In-memory loader for the UNMODIFIED reference (Python 2 sources under /root/reference).

TEST INFRASTRUCTURE ONLY.  Nothing under ``pyhillfit_b200/`` may import this
module.  It exists only in the build container (``/root/reference`` is absent on
the GPU box); its sole consumers are ``oracle/gen_golden.py`` (which writes the
committed fixtures under ``tests/golden/``) and CPU tests that are skipped when
``/root/reference`` is missing.

No reference source is copied into this repository: the files are read where
they lie, passed through a two-rule Python-2 -> Python-3 text shim
(``print x`` -> ``print(x)``; ``raw_input`` -> ``input``; ``xrange`` -> ``range``)
and ``exec``-ed into fresh module namespaces.  For the two script files
(PyHillFit.py / PyHillTemp.py), which run argparse + ``import cma`` + matplotlib at
import time, only the named ``FunctionDef`` nodes are kept (AST extraction) and the
Python-2 integer divisions inside ``do_mcmc`` are rewritten to ``//``.
"""
import ast
import os
import re
import sys
import types

REF_ROOT = os.environ.get("PHF_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "python", "doseresponse.py"))


def _py2to3(src):
    src = re.sub(r'^(\s*)print (.*)$', r'\1print(\2)', src, flags=re.M)
    src = src.replace("raw_input(", "input(").replace("xrange(", "range(")
    src = re.sub(r'except (\w+)\s*,\s*(\w+):', r'except \1 as \2:', src)   # py2 "except Exception,e:"
    return src


def load_doseresponse():
    """python/doseresponse.py as a live module (all functions run unmodified)."""
    path = os.path.join(REF_ROOT, "python", "doseresponse.py")
    with open(path) as f:
        src = _py2to3(f.read())
    mod = types.ModuleType("ref_doseresponse")
    mod.__file__ = path
    mod.__dict__["sys"] = sys  # the reference forgets to import sys (doseresponse.py:315)
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def _extract_functions(path, names, int_div=False):
    with open(path) as f:
        src = _py2to3(f.read())
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(keep) == len(names), (names, [k.name for k in keep])
    if int_div:
        class _FloorDiv(ast.NodeTransformer):
            # Python-2 `/` on ints (PyHillTemp.py:70-71,109,112) -> `//`.  Only integer
            # operands are affected by py2 semantics; the float expressions in do_mcmc
            # ((t-1.)/t, 1./t, 1./(s+1.)**0.6) carry an explicit float literal.
            def visit_BinOp(self, node):
                self.generic_visit(node)
                if isinstance(node.op, ast.Div):
                    has_float = any(isinstance(c, ast.Constant) and isinstance(c.value, float)
                                    for c in ast.walk(node))
                    if not has_float:
                        node.op = ast.FloorDiv()
                return node
        keep = [_FloorDiv().visit(k) for k in keep]
    module = ast.Module(body=keep, type_ignores=[])
    ast.fix_missing_locations(module)
    return compile(module, path, "exec")


def load_hierarchical_functions(dr):
    """The four live hierarchical target functions of python/PyHillFit.py:113-193."""
    import numpy as np
    import scipy.stats as st
    ns = {"dr": dr, "np": np, "st": st, "sys": sys, "pic50_prior": [-2.]}  # PyHillFit.py:214-215
    code = _extract_functions(
        os.path.join(REF_ROOT, "python", "PyHillFit.py"),
        ["log_data_likelihood", "log_hill_i_log_logistic_likelihood",
         "log_pic50_i_logistic_likelihood", "log_target_distribution"])
    exec(code, ns)
    return ns


def load_do_mcmc(dr, namespace):
    """python/PyHillTemp.py:57-125 `do_mcmc`, run against caller-supplied globals
    (args, responses, where_r_0, where_r_100, where_r_other, concs, pi_bit, num_params)."""
    import numpy as np
    import numpy.random as npr
    ns = {"dr": dr, "np": np, "npr": npr, "sys": sys}
    ns.update(namespace)
    code = _extract_functions(os.path.join(REF_ROOT, "python", "PyHillTemp.py"), ["do_mcmc"], int_div=True)
    exec(code, ns)
    return ns["do_mcmc"], ns


def load_construct_cdfs():
    """python/construct_hierarchical_cdfs.py:32-58 `construct_posterior_predictive_cdfs` (the script itself runs
    argparse at import, so only the function is extracted)."""
    import numpy as np
    import scipy.stats as st
    ns = {"np": np, "st": st, "xrange": range}
    code = _extract_functions(os.path.join(REF_ROOT, "python", "construct_hierarchical_cdfs.py"),
                              ["construct_posterior_predictive_cdfs"])
    exec(code, ns)
    return ns["construct_posterior_predictive_cdfs"]
