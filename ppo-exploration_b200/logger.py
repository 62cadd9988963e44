"""Diagnostics logger with the reference's call surface and CSV schema (logger.py of the reference, :13-52 CSV format,
:126-141 Logger.record/dump, :203-246 module-level record/dump/configure) so that `Visualisation.ipynb` and any
pandas-based tooling keep working on logs written by ppx learners.

    from ppo_exploration_b200 import logger
    logger.configure("PPO", "Swimmer-v2", log_to_file=True)
    m = ppx.PPO(env=env, logger=logger, ...)      # the learners call logger.record(key, value) / logger.dump(step)

CSV schema (what the reference writes): one row per dump(); a key "group/name" is stored under column "name" (the SECOND
path component, reference :27-29); the header grows when new keys appear -- the file is then rewritten with the longer
header and the earlier rows padded with empty cells; values are written with str(); a missing value is an empty cell.
Host I/O only -- nothing here touches the GPU.
"""
import datetime
import os
import sys
from collections import OrderedDict


def _column(key):
    """'train/value_loss' -> 'value_loss' (reference logger.py:27-29: the part after the first '/', when there is one
    past position 0)."""
    return key.split('/')[1] if key.find('/') > 0 else key


class CSVOutputFormat(object):
    def __init__(self, filename):
        self.filename = filename
        self.keys = []
        self.rows = []                      # every row written so far, as {column: str(value)}
        self.file = open(filename, "w+t")

    def _line(self, row):
        return ",".join(row.get(k, "") for k in self.keys) + "\n"

    def write(self, key_values):
        row = OrderedDict()
        for key in sorted(key_values.keys()):
            value = key_values[key]
            row[_column(key)] = "" if value is None else str(value)
        new = [k for k in row if k not in self.keys]
        if new:                             # longer header: rewrite the file, earlier rows get empty cells
            self.keys.extend(sorted(new))
            self.file.seek(0)
            self.file.truncate()
            self.file.write(",".join(self.keys) + "\n")
            for old in self.rows:
                self.file.write(self._line(old))
        self.rows.append(row)
        self.file.write(self._line(row))
        self.file.flush()

    def close(self):
        self.file.close()


class HumanOutputFormat(object):
    """Boxed key/value table on a stream, grouped by the key's tag (reference :60-118)."""

    def __init__(self, filename_or_file):
        self.own_file = isinstance(filename_or_file, str)
        self.file = open(filename_or_file, "wt") if self.own_file else filename_or_file

    @staticmethod
    def _short(text, width=23):
        return text if len(text) <= width else text[:width - 3] + "..."

    def write(self, key_values):
        cells, tag = OrderedDict(), None
        for key, value in sorted(key_values.items()):
            text = f"{value:<8.3g}" if isinstance(value, float) else str(value)
            if key.find("/") > 0:
                tag = key[:key.find("/") + 1]
                cells[self._short(tag)] = ""
            if tag is not None and tag in key:
                key = "   " + key[len(tag):]
            cells[self._short(key)] = self._short(text)
        if not cells:
            return
        kw, vw = max(map(len, cells)), max(map(len, cells.values()))
        rule = "-" * (kw + vw + 7)
        out = [rule] + [f"| {k}{' ' * (kw - len(k))} | {v}{' ' * (vw - len(v))} |" for k, v in cells.items()] + [rule]
        self.file.write("\n".join(out) + "\n")
        self.file.flush()

    def close(self):
        if self.own_file:
            self.file.close()


class Logger(object):
    CURRENT = None

    def __init__(self, outputs, folder='./logs'):
        self.name_to_value = {}
        self.dir = folder
        self.outputs = list(outputs) if isinstance(outputs, (list, tuple)) else [outputs]

    def record(self, key, value):
        self.name_to_value[key] = value

    def dump(self, step=0):
        for writer in self.outputs:
            writer.write(dict(self.name_to_value))
        self.name_to_value.clear()

    def get_dir(self):
        return self.dir

    def close(self):
        for writer in self.outputs:
            writer.close()


Logger.CURRENT = Logger([HumanOutputFormat(sys.stdout)])


def record(key, value):
    Logger.CURRENT.record(key, value)


def dump(step=0):
    Logger.CURRENT.dump(step)


def configure(algorithm, environment, log_to_file=False, folder=None):
    """Same arguments and directory layout as the reference (:222-246): <folder>/<algorithm>/<environment>/run-<time>.csv"""
    folder = os.path.join("./logs" if folder is None else folder, algorithm, environment)
    outputs = [HumanOutputFormat(sys.stdout)]
    if log_to_file:
        os.makedirs(folder, exist_ok=True)
        name = "run" + datetime.datetime.now().strftime("-%Y-%m-%d-%H-%M-%S-%f") + ".csv"
        outputs.append(CSVOutputFormat(os.path.join(folder, name)))
    Logger.CURRENT = Logger(outputs, folder=folder)
    print(f"Logging to {folder}")
    return Logger.CURRENT
