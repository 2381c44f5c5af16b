"""Case input: a reference ``case.py`` / ``.json`` / ``simulation.inp`` -> the case dictionary.

Restates the input side of the reference toolchain and of the simulation start-up:
  * ``toolchain/mfc/run/input.py:166-200`` (``load``): a ``.py`` case is *executed* with
    ``--mfc '<json of the run arguments>'`` (+ user arguments after ``--``) and prints one JSON
    dictionary; ``.json`` is read as is;
  * ``toolchain/mfc/case.py:79-115`` (``get_inp``): the dictionary becomes the Fortran namelist
    ``&user_inputs ... &end/`` of ``<target>.inp`` (strings of length one -- the logicals ``T``/``F`` --
    unquoted, longer strings quoted);
  * ``src/simulation/m_start_up.fpp:81-175`` (``s_read_input_file``): ``simulation.inp`` is read back
    by ``read(1, NML=user_inputs)``.
Only what the 5-equation hot path consumes is interpreted (``params.params_from_case``); every other
key is carried through untouched, and unsupported physics is rejected there like the reference's
``@:PROHIBIT``.
"""
import json
import os
import re
import subprocess
import sys


class CaseError(ValueError):
    pass


def load_case(path, args=(), mfc_args=None):
    """Case dictionary of ``path`` (``.py`` executed as the reference does, ``.json``, or a ``.inp`` namelist)."""
    if not os.path.exists(path):
        raise CaseError(f"Input file '{path}' does not exist. Please check the path is valid.")
    if path.endswith(".py"):
        # the reference hands its run arguments over (toolchain/mfc/state.py ARGS()); the case files that look at them
        # read "gpu" (benchmarks/*/case.py:22) -- this library always runs on the GPU
        margs = {"gpu": "acc", "mpi": True, "targets": ["pre_process", "simulation"], "case_optimization": True}
        margs.update(mfc_args or {})
        cmd = [sys.executable, os.path.abspath(path), "--mfc", json.dumps(margs)] + list(args)
        r = subprocess.run(cmd, capture_output=True, text=True, cwd=os.path.dirname(os.path.abspath(path)) or ".")
        if r.returncode != 0:
            raise CaseError(f"Input file {path} terminated with a non-zero exit code. Please make sure running the file "
                            f"doesn't produce any errors.\n{r.stderr[-2000:]}")
        try:
            case = json.loads(r.stdout)
        except ValueError as e:
            raise CaseError(f"Input file {path} did not print a JSON dictionary: {e}") from e
    elif path.endswith(".json"):
        with open(path) as f:
            case = json.load(f)
    elif path.endswith(".inp"):
        case = read_namelist(path)
    else:
        raise CaseError("Unrecognized input file format. Please check the extension. [.py, .json, .inp]")
    if not isinstance(case, dict):
        raise CaseError("the case must be one dictionary of parameters")
    return case


def _fmt_value(val):
    # toolchain/mfc/case.py:100-103
    if not isinstance(val, str) or len(val) == 1:
        return f"{val}"
    return f"'{val}'"


def write_namelist(path, case, keys=None):
    """``<target>.inp`` exactly as ``get_inp`` lays it out; ``keys`` filters like MASTER_KEYS."""
    lines = ["&user_inputs"]
    for k, v in case.items():
        if keys is not None and k not in keys:
            continue
        lines.append(f"{k} = {_fmt_value(v)}")
    lines.append("&end/")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


_NUM = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)([eEdD][+-]?\d+)?$")


def _parse_value(s):
    s = s.strip().rstrip(",").strip()
    if len(s) >= 2 and s[0] == s[-1] and s[0] in "'\"":
        return s[1:-1]
    u = s.upper()
    if u in ("T", ".TRUE.", "TRUE"):
        return "T"
    if u in ("F", ".FALSE.", "FALSE"):
        return "F"
    if _NUM.match(s):
        if re.match(r"^[+-]?\d+$", s):
            return int(s)
        return float(s.replace("d", "e").replace("D", "e"))
    return s


def read_namelist(path):
    """Parse the ``&user_inputs`` group of a reference ``.inp`` file (one ``key = value`` per line)."""
    case, inside = {}, False
    with open(path) as f:
        for raw in f:
            line = raw.split("!")[0].strip()
            if not line:
                continue
            if line.lower().startswith("&user_inputs"):
                inside = True
                continue
            if line.startswith("&end") or line == "/":
                break
            if not inside or "=" not in line:
                continue
            k, v = line.split("=", 1)
            case[k.strip()] = _parse_value(v)
    if not inside:
        raise CaseError(f"{path}: no &user_inputs namelist group")
    return case
