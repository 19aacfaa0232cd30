from openpvsg_b200.relation_set import *  # noqa: F401,F403  (same function names as utils/relation_matching.py)
