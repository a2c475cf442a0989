"""Experiment-file helpers (counterparts of /root/reference/src/utilities/utils.py:19-93,149-168): nested update of
the base config, expansion of a grid section into its experiments, a flushing file logger.  MLflow bookkeeping
(utils.py:96-146) is not rebuilt: the driver writes plain JSON next to the predictions instead."""
import collections.abc
import itertools
import logging
import os


def nested_dict_update(d, u):
    """d updated in place with u, descending into nested mappings (utils.py:19-31)"""
    for key, value in u.items():
        if isinstance(value, collections.abc.Mapping):
            d[key] = nested_dict_update(d.get(key, {}), value)
        else:
            d[key] = value
    return d


def _leaves(node, path=()):
    """(key path, list of values) for every list leaf of a nested dict-of-lists, in file order"""
    for key, value in node.items():
        if isinstance(value, collections.abc.Mapping):
            yield from _leaves(value, path + (key,))
        elif isinstance(value, list):
            yield path + (key,), value
        else:
            raise ValueError("Only dict or lists!!!")


def make_grid(dict_of_list):
    """Every combination of the listed values as a nested dict (utils.py:81-93).  Order: the cartesian product
    with the LAST listed key varying fastest, as itertools.product over the leaves in file order."""
    leaves = list(_leaves(dict_of_list))
    out = []
    for combo in itertools.product(*[values for _, values in leaves]):
        exp = {}
        for (path, _), value in zip(leaves, combo):
            node = exp
            for key in path[:-1]:
                node = node.setdefault(key, {})
            node[path[-1]] = value
        out.append(exp)
    return out


class FlushFileHandler(logging.FileHandler):
    def emit(self, record):
        super().emit(record)
        self.flush()


def get_experiment_logger(dest):
    """logger writing to <dest>/log.txt (utils.py:149-168)"""
    logger = logging.getLogger("cbrs.experiment." + os.path.abspath(dest))
    logger.setLevel(logging.INFO)
    logger.propagate = False
    for handler in list(logger.handlers):
        logger.removeHandler(handler)
    handler = FlushFileHandler(os.path.join(dest, "log.txt"))
    handler.setFormatter(logging.Formatter("%(asctime)s - %(levelname)s - %(message)s"))
    logger.addHandler(handler)
    return logger
