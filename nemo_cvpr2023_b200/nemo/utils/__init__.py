from .misc_utils import *  # noqa
