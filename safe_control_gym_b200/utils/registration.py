'''String-id registry: the plugin API of the reference, kept verbatim in behaviour.

Mirrors safe_control_gym/utils/registration.py:118-135 (`register`, `make`, `spec`, `get_config`):
ids map to a 'module:Class' entry point plus a packaged YAML with the default kwargs; `Spec.make`
attaches `obj.spec` (or `obj.instance_spec` when the object already has a `spec`).
'''

import copy
import importlib
import importlib.resources as pkg_resources

import yaml


def load(name):
    '''`module_path:callable_name` -> callable (registration.py:18-23).'''
    mod_name, attr_name = name.split(':')
    return getattr(importlib.import_module(mod_name), attr_name)


class Spec:
    def __init__(self, idx, entry_point=None, config_entry_point=None):
        self.idx = idx
        self.entry_point = entry_point
        self.config_entry_point = config_entry_point

    def __repr__(self):
        return f'Spec({self.idx})'

    def get_config(self):
        cep = self.config_entry_point
        if cep is None:
            return {}
        if not isinstance(cep, str):
            raise Exception(f'Config type {cep} is not supported.')
        if cep.endswith('.yaml'):
            mod_name, config_name = cep.split(':')
            with pkg_resources.files(mod_name).joinpath(config_name).open('r') as f:
                return yaml.load(f, Loader=yaml.FullLoader)
        return load(cep)

    def make(self, *args, **kwargs):
        if self.entry_point is None:
            raise Exception(f'Attempting to make deprecated env {self.idx}.')
        factory = self.entry_point if callable(self.entry_point) else load(self.entry_point)
        obj = factory(*args, **kwargs)
        spec = copy.deepcopy(self)
        if hasattr(obj, 'spec'):
            obj.instance_spec = spec
        else:
            obj.spec = spec
        return obj


class Registry:
    def __init__(self):
        self.specs = {}

    def make(self, path, *args, **kwargs):
        return self.specs[path].make(*args, **kwargs)

    def all(self):
        return self.specs.values()

    def spec(self, path):
        if ':' in path:
            mod_name, idx = path.split(':')
            try:
                importlib.import_module(mod_name)
            except BaseException:
                raise Exception(f'A module ({mod_name}) was specified for the environment but was not found.')
        else:
            idx = path
        try:
            return self.specs[idx]
        except BaseException:
            raise Exception('Key not found in registry.')

    def register(self, idx, **kwargs):
        if idx in self.specs:
            raise Exception(f'Cannot re-register id: {idx}')
        self.specs[idx] = Spec(idx, **kwargs)


registry = Registry()


def register(idx, **kwargs):
    return registry.register(idx, **kwargs)


def make(idx, *args, **kwargs):
    return registry.make(idx, *args, **kwargs)


def spec(idx):
    return registry.spec(idx)


def get_config(idx):
    return registry.spec(idx).get_config()
