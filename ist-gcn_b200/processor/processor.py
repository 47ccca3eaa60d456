"""``processor.processor.Processor`` (reference: processor/processor.py) lives in recognition.py."""
from .recognition import Processor  # noqa: F401
