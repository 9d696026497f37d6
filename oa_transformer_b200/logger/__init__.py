"""setup_logging of OATrans/logger/logger.py: console + rotating file handler under the run's log dir."""
import logging
import logging.handlers
from pathlib import Path


def setup_logging(save_dir, default_level=logging.INFO):
    save_dir = Path(save_dir)
    root = logging.getLogger()
    root.setLevel(default_level)
    if not any(isinstance(h, logging.StreamHandler) for h in root.handlers):
        ch = logging.StreamHandler()
        ch.setFormatter(logging.Formatter("%(message)s"))
        root.addHandler(ch)
    fh = logging.handlers.RotatingFileHandler(str(save_dir / "info.log"), maxBytes=10485760, backupCount=20,
                                              encoding="utf8")
    fh.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s"))
    root.addHandler(fh)
