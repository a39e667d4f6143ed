"""On-disk formats of a training run: progress.csv, variant.json, debug.log and the
params.pkl / itr_N.pkl snapshots, laid out like the reference's rllab-style logger writes them
(ast_sac/core/logging.py:84-336, ast_sac/launchers/launcher_utils.py:190-300), so that tools reading a reference run
directory read ours.

What is reproduced, rule by rule:
  * ``record_tabular`` / ``record_dict(d, prefix)`` collect (key, value) pairs of the epoch; ``dump_tabular`` writes
    one csv row.  The column set is the SORTED key set of the first dump and stays fixed: later extra keys are ignored
    with a warning (``csv.DictWriter(extrasaction="ignore")``), the header is written once (logging.py:274-312).
  * ``save_itr_params(itr, params)`` with snapshot modes 'all' (itr_N.pkl), 'last' (params.pkl), 'gap', 'gap_and_last',
    'none', written with ``torch.save`` (logging.py:314-336).
  * ``log_variant`` writes the variant dict as sorted, indented JSON (logging.py:249-252); ``log`` prefixes lines with
    a timestamp and the experiment name and appends them to the text outputs (logging.py:174-188).
  * ``setup_logger`` creates ``<base>/<exp-prefix>/<exp-prefix>_<timestamp>_<id>--s-<seed>/`` and wires the four files.
The column names themselves come from the diagnostics of the trainer / collectors / replay buffer (``record_dict`` with
the reference's prefixes, rl_algorithm.py:76-141); tests/golden/progress_format.json holds the column set, snapshot
keys and file list of a run of the UNMODIFIED reference, and tests/test_rl_cpu.py compares against it.
"""
from __future__ import annotations

import csv
import datetime
import json
import os
import os.path as osp
import sys
from collections import OrderedDict
from enum import Enum


class _Encoder(json.JSONEncoder):
    """logging.py:59-72: types, enums and callables by name."""

    def default(self, o):
        if isinstance(o, type):
            return {'$class': o.__module__ + "." + o.__name__}
        if isinstance(o, Enum):
            return {'$enum': o.__module__ + "." + o.__class__.__name__ + '.' + o.name}
        if callable(o):
            return {'$function': o.__module__ + "." + o.__name__}
        return json.JSONEncoder.default(self, o)


class Logger:
    def __init__(self):
        self._prefixes, self._prefix_str = [], ''
        self._tabular_prefixes, self._tabular_prefix_str = [], ''
        self._tabular = []
        self._text_fds, self._tabular_fds = {}, {}
        self._tabular_header_written = set()
        self._tabular_keys = {}
        self._snapshot_dir = None
        self._snapshot_mode = 'all'
        self._snapshot_gap = 1
        self._log_tabular_only = False
        self._print = True

    # -- outputs -------------------------------------------------------------------------------------
    @staticmethod
    def _open(file_name, fds, mode='a'):
        if file_name not in fds:
            os.makedirs(osp.dirname(osp.abspath(file_name)), exist_ok=True)
            fds[file_name] = open(file_name, mode)

    def add_text_output(self, file_name):
        self._open(file_name, self._text_fds, 'a')

    def add_tabular_output(self, file_name, relative_to_snapshot_dir=False):
        if relative_to_snapshot_dir:
            file_name = osp.join(self._snapshot_dir, file_name)
        self._open(file_name, self._tabular_fds, 'w')

    def remove_text_output(self, file_name):
        if file_name in self._text_fds:
            self._text_fds.pop(file_name).close()

    def remove_tabular_output(self, file_name, relative_to_snapshot_dir=False):
        if relative_to_snapshot_dir:
            file_name = osp.join(self._snapshot_dir, file_name)
        fd = self._tabular_fds.pop(file_name, None)
        if fd is not None:
            self._tabular_header_written.discard(fd)
            fd.close()

    def close(self):
        for fd in list(self._text_fds.values()) + list(self._tabular_fds.values()):
            fd.close()
        self._text_fds, self._tabular_fds = {}, {}
        self._tabular_header_written = set()
        self._tabular_keys = {}

    def set_snapshot_dir(self, dir_name):
        self._snapshot_dir = dir_name

    def get_snapshot_dir(self):
        return self._snapshot_dir

    def set_snapshot_mode(self, mode):
        self._snapshot_mode = mode

    def get_snapshot_mode(self):
        return self._snapshot_mode

    def set_snapshot_gap(self, gap):
        self._snapshot_gap = gap

    def get_snapshot_gap(self):
        return self._snapshot_gap

    def set_log_tabular_only(self, log_tabular_only):
        self._log_tabular_only = log_tabular_only

    def set_print(self, on: bool):
        """(not in the reference) keep the text / csv files but do not echo to stdout."""
        self._print = on

    # -- text ----------------------------------------------------------------------------------------
    def push_prefix(self, prefix):
        self._prefixes.append(prefix)
        self._prefix_str = ''.join(self._prefixes)

    def pop_prefix(self):
        del self._prefixes[-1]
        self._prefix_str = ''.join(self._prefixes)

    def log(self, s, with_prefix=True, with_timestamp=True):
        out = s
        if with_prefix:
            out = self._prefix_str + out
        if with_timestamp:
            out = "%s | %s" % (datetime.datetime.now().strftime('%Y-%m-%d %H:%M:%S.%f %Z'), out)
        if not self._log_tabular_only:
            if self._print:
                print(out)
            for fd in self._text_fds.values():
                fd.write(out + '\n')
                fd.flush()
            sys.stdout.flush()

    def log_variant(self, log_file, variant_data):
        os.makedirs(osp.dirname(osp.abspath(log_file)), exist_ok=True)
        with open(log_file, "w") as f:
            json.dump(variant_data, f, indent=2, sort_keys=True, cls=_Encoder)

    # -- table ---------------------------------------------------------------------------------------
    def push_tabular_prefix(self, key):
        self._tabular_prefixes.append(key)
        self._tabular_prefix_str = ''.join(self._tabular_prefixes)

    def pop_tabular_prefix(self):
        del self._tabular_prefixes[-1]
        self._tabular_prefix_str = ''.join(self._tabular_prefixes)

    def record_tabular(self, key, val):
        self._tabular.append((self._tabular_prefix_str + str(key), str(val)))

    def record_dict(self, d, prefix=None):
        if prefix is not None:
            self.push_tabular_prefix(prefix)
        for k, v in d.items():
            self.record_tabular(k, v)
        if prefix is not None:
            self.pop_tabular_prefix()

    def get_table_dict(self):
        return dict(self._tabular)

    def get_table_key_set(self):
        return set(key for key, value in self._tabular)

    def dump_tabular(self, *args, **kwargs):
        wh = kwargs.pop("write_header", None)
        if len(self._tabular) > 0:
            if not self._log_tabular_only:
                width = max(len(k) for k, _ in self._tabular)
                rule = '-' * (width + 2 + max(len(v) for _, v in self._tabular))
                for line in [rule] + ["%-*s  %s" % (width, k, v) for k, v in self._tabular] + [rule]:
                    self.log(line, *args, **kwargs)
            tabular_dict = dict(self._tabular)
            for filename, fd in list(self._tabular_fds.items()):
                keys = self._tabular_keys.get(filename)
                if keys is None:                  # the first dump fixes the columns (logging.py:287-291)
                    keys = list(sorted(tabular_dict.keys()))
                    self._tabular_keys[filename] = keys
                elif set(keys) != set(tabular_dict.keys()):
                    print("Warning: CSV key mismatch")
                    print("extra keys in 0th iter", set(keys) - set(tabular_dict.keys()))
                    print("extra keys in cur iter", set(tabular_dict.keys()) - set(keys))
                writer = csv.DictWriter(fd, fieldnames=keys, extrasaction="ignore")
                if wh or (wh is None and fd not in self._tabular_header_written):
                    writer.writeheader()
                    self._tabular_header_written.add(fd)
                writer.writerow(tabular_dict)
                fd.flush()
            del self._tabular[:]

    # -- snapshots -----------------------------------------------------------------------------------
    def save_itr_params(self, itr, params):
        import torch
        if not self._snapshot_dir:
            return
        mode = self._snapshot_mode
        if mode == 'all':
            torch.save(params, osp.join(self._snapshot_dir, 'itr_%d.pkl' % itr))
        elif mode == 'last':
            torch.save(params, osp.join(self._snapshot_dir, 'params.pkl'))
        elif mode == 'gap':
            if itr % self._snapshot_gap == 0:
                torch.save(params, osp.join(self._snapshot_dir, 'itr_%d.pkl' % itr))
        elif mode == 'gap_and_last':
            if itr % self._snapshot_gap == 0:
                torch.save(params, osp.join(self._snapshot_dir, 'itr_%d.pkl' % itr))
            torch.save(params, osp.join(self._snapshot_dir, 'params.pkl'))
        elif mode == 'none':
            pass
        else:
            raise NotImplementedError(mode)


logger = Logger()


def create_exp_name(exp_prefix, exp_id=0, seed=0):
    """launcher_utils.py:181-195."""
    return "%s_%s_%04d--s-%d" % (exp_prefix, datetime.datetime.now().strftime('%Y_%m_%d_%H_%M_%S'), exp_id, seed)


def create_log_dir(exp_prefix, exp_id=0, seed=0, base_log_dir=None, include_exp_prefix_sub_dir=True):
    """launcher_utils.py:198-225 (the reference's default base directory is <repo>/run/logs)."""
    exp_name = create_exp_name(exp_prefix, exp_id=exp_id, seed=seed)
    if base_log_dir is None:
        base_log_dir = osp.join(os.getcwd(), "run", "logs")
    if include_exp_prefix_sub_dir:
        log_dir = osp.join(base_log_dir, exp_prefix.replace("_", "-"), exp_name)
    else:
        log_dir = osp.join(base_log_dir, exp_name)
    if osp.exists(log_dir):
        print("WARNING: Log directory already exists {}".format(log_dir))
    os.makedirs(log_dir, exist_ok=True)
    return log_dir


def setup_logger(exp_prefix="default", variant=None, text_log_file="debug.log", variant_log_file="variant.json",
                 tabular_log_file="progress.csv", snapshot_mode="last", snapshot_gap=1, log_tabular_only=False,
                 log_dir=None, target: Logger = None, **create_log_dir_kwargs):
    """launcher_utils.py:227-300 without its git bookkeeping: returns the run directory."""
    lg = target or logger
    first_time = log_dir is None
    if first_time:
        log_dir = create_log_dir(exp_prefix, **create_log_dir_kwargs)
    if variant is not None:
        lg.log("Variant:")
        lg.log(json.dumps(variant, indent=2, cls=_Encoder, default=str))
        lg.log_variant(osp.join(log_dir, variant_log_file), variant)
    lg.add_text_output(osp.join(log_dir, text_log_file))
    lg.add_tabular_output(osp.join(log_dir, tabular_log_file))
    lg.set_snapshot_dir(log_dir)
    lg.set_snapshot_mode(snapshot_mode)
    lg.set_snapshot_gap(snapshot_gap)
    lg.set_log_tabular_only(log_tabular_only)
    lg.push_prefix("[%s] " % log_dir.split("/")[-1])
    return log_dir


def create_stats_ordered_dict(name, data, stat_prefix=None):
    """eval_util.py:74-118 for tensors / arrays: Mean, Std (population), Max, Min."""
    import numpy as np
    import torch
    if stat_prefix is not None:
        name = "{}{}".format(stat_prefix, name)
    if torch.is_tensor(data):
        data = data.detach().to(torch.float64).cpu().numpy()
    data = np.asarray(data, dtype=np.float64)
    if data.size == 0:
        return OrderedDict()
    return OrderedDict([(name + ' Mean', float(np.mean(data))), (name + ' Std', float(np.std(data))),
                        (name + ' Max', float(np.max(data))), (name + ' Min', float(np.min(data)))])
