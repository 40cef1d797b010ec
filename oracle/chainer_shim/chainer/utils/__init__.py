from . import argument
from . import type_check
