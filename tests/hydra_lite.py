"""TEST INFRASTRUCTURE -- a minimal composer for the reference's hydra config tree (/root/reference/configs), enough to
produce the resolved config of one experiment without hydra-core / omegaconf (not installable offline).

Implements exactly the hydra 1.1 features those files use: `defaults` lists with `group@package: option`, absolute
(`/group/...`) and group-relative entries, `override` entries (replace the option chosen by an earlier default for the
same group@package), `_self_` ordering, `# @package _global_` headers, empty options (`group@pkg:` = mandatory, chosen
by an override), and `${a.b}` / `${.a}` / `${..a}` interpolations.  Later entries win, dicts merge recursively.
"""
import os
import re

import yaml


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v
    return dst


def _nest(package, value):
    for part in reversed([p for p in package.split(".") if p]):
        value = {part: value}
    return value


def _load(root, path):
    with open(os.path.join(root, path + ".yaml")) as f:
        text = f.read()
    body = yaml.safe_load(text) or {}
    global_pkg = bool(re.match(r"\s*#\s*@package\s+_global_", text))
    return body, global_pkg


def _parse_entry(e):
    """-> (override, group, package or None, option or None)"""
    if isinstance(e, str):
        if e == "_self_":
            return None
        key, opt = e, None
        if "@" not in e:  # plain config file in the same group (e.g. "base_visualizer", "dd3d_kitti_dla34")
            return (False, None, None, e)
    else:
        (key, opt), = e.items()
    override = key.startswith("override ")
    if override:
        key = key[len("override "):]
    group, _, pkg = key.partition("@")
    return (override, group, pkg if "@" in key else None, opt)


def _choice(choices, gpath, pkg):
    """Option an `override` entry picked for this group: exact group@package match, else a unique match on the group."""
    if (gpath, pkg) in choices:
        return choices[(gpath, pkg)]
    hits = [v for (g, _), v in choices.items() if g == gpath]
    return hits[0] if len(hits) == 1 else None


def _compose(root, path, package, choices):
    """Config at `path` (relative to root, no extension) placed at `package` ('' = global)."""
    body, global_pkg = _load(root, path)
    if global_pkg:
        package = ""
    defaults = body.pop("defaults", [])
    here = os.path.dirname(path)
    out = {}
    self_done = False
    for e in defaults:
        pe = _parse_entry(e)
        if pe is None:
            _merge(out, _nest(package, body))
            self_done = True
            continue
        override, group, pkg, opt = pe
        if override:
            continue  # collected by collect_overrides before composing
        if group is None:  # sibling config file
            _merge(out, _compose(root, os.path.join(here, opt), package, choices))
            continue
        absolute = group.startswith("/")
        gpath = group.lstrip("/") if absolute else os.path.join(here, group)
        if opt is None:
            # "common/test@TEST"-style entries name a file directly; "group@pkg:" entries need a choice
            if os.path.exists(os.path.join(root, gpath + ".yaml")):
                target = gpath
            else:
                chosen = _choice(choices, gpath, pkg)
                if chosen is None:
                    raise KeyError(f"no option chosen for {gpath}@{pkg}")
                target = os.path.join(gpath, chosen)
        else:
            target = os.path.join(gpath, _choice(choices, gpath, pkg) or opt)
        if pkg is None:
            sub_pkg = ".".join(p for p in (package, os.path.basename(gpath)) if p)
        elif pkg == "":
            sub_pkg = package
        elif absolute or global_pkg or not package:
            sub_pkg = pkg
        else:
            sub_pkg = package + "." + pkg
        _merge(out, _compose(root, target, sub_pkg, choices))
    if not self_done:
        _merge(out, _nest(package, body))
    return out


def _collect_overrides(root, path, choices):
    body, _ = _load(root, path)
    here = os.path.dirname(path)
    for e in body.get("defaults", []):
        pe = _parse_entry(e)
        if pe is None:
            continue
        override, group, pkg, opt = pe
        if group is None:
            _collect_overrides(root, os.path.join(here, opt), choices)
        elif override:
            choices[(group.lstrip("/"), pkg)] = opt


def _resolve(cfg):
    pat = re.compile(r"^\$\{([^}]+)\}$")

    def get(path_parts):
        node = cfg
        for p in path_parts:
            node = node[p]
        return node

    def walk(node, trail):
        for k, v in (node.items() if isinstance(node, dict) else enumerate(node)):
            if isinstance(v, (dict, list)):
                walk(v, trail + [k])
            elif isinstance(v, str):
                m = pat.match(v)
                if not m:
                    continue
                ref = m.group(1)
                if ref.startswith("."):
                    dots = len(ref) - len(ref.lstrip("."))
                    base = trail[:len(trail) - (dots - 1)]
                    parts = base + ref.lstrip(".").split(".")
                else:
                    parts = ref.split(".")
                val = get(parts)
                if isinstance(val, str) and pat.match(val):
                    continue  # resolved on a later pass
                node[k] = val

    for _ in range(4):
        walk(cfg, [])
    return cfg


def compose_experiment(root, experiment):
    """Resolved config of `+experiments=<experiment>` on top of configs/defaults.yaml (scripts/train.py's config_name)."""
    choices = {}
    _collect_overrides(root, os.path.join("experiments", experiment), choices)
    cfg = _compose(root, "defaults", "", choices)
    _merge(cfg, _compose(root, os.path.join("experiments", experiment), "", choices))
    cfg.pop("hydra", None)
    return _resolve(cfg)
