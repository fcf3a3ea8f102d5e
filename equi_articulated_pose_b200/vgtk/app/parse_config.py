import argparse
import json


class HierarchyArgmentParser():
    """Groups of flags; every group scans the whole argv with parse_known_args
    (reference: vgtk/vgtk/app/parse_config.py:7-29).  Groups named experiment/train/eval/test are
    flattened into the top-level namespace, the others are attached as attributes."""

    FLAT = ('experiment', 'train', 'eval', 'test')

    def __init__(self, flat_groups=None):
        self.parsers = {}
        self.flat = tuple(flat_groups) if flat_groups is not None else self.FLAT

    def add_parser(self, name):
        p = argparse.ArgumentParser(add_help=False)
        self.parsers[name] = p
        return p

    def parse_args(self, argv=None):
        opt = argparse.Namespace()
        for name, p in self.parsers.items():
            ns, _ = p.parse_known_args(argv)
            if name in self.flat:
                for k, v in vars(ns).items():
                    setattr(opt, k, v)
            else:
                setattr(opt, name, ns)
        return opt


def dump_args(opt, path=None):
    def enc(ns):
        return {k: (enc(v) if isinstance(v, argparse.Namespace) else v) for k, v in vars(ns).items()}
    txt = json.dumps(enc(opt), indent=2, default=str)
    if path is not None:
        with open(path, 'w') as fh:
            fh.write(txt)
    return txt
