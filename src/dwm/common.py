"""JSON `_class_name` object factory — the plug-in boundary of OpenDWM
(reference: src/dwm/common.py:133-186; schema in configs/README.md:51-70).

Any dict holding "_class_name" is instantiated recursively; the special name
"get_class" returns the class object itself; extra kwargs are injected only at the
top level; `global_state` is a process-wide registry read with `get_state`.
"""
import importlib

global_state = {}


def get_class(class_name: str):
    if "." in class_name:
        module_name, _, attr = class_name.rpartition(".")
        return getattr(importlib.import_module(module_name), attr)
    if class_name in globals():
        return globals()[class_name]
    raise RuntimeError("Failed to find the class {}.".format(class_name))


def create_instance(class_name: str, **kwargs):
    return get_class(class_name)(**kwargs)


def instantiate_config(_config: dict, level: int = 0):
    return {
        k: create_instance_from_config(v, level + 1)
        for k, v in _config.items() if k != "_class_name"
    }


def create_instance_from_config(_config, level: int = 0, **kwargs):
    if isinstance(_config, dict):
        if "_class_name" not in _config:
            return instantiate_config(_config, level)
        args = instantiate_config(_config, level)
        if level == 0:
            args.update(kwargs)
        if _config["_class_name"] == "get_class":
            return get_class(**args)
        return create_instance(_config["_class_name"], **args)
    if isinstance(_config, list):
        return [create_instance_from_config(i, level + 1) for i in _config]
    return _config


def get_state(key: str):
    return global_state[key]
