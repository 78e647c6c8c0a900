"""``mad.structure_utils`` of the reference (mad/structure_utils.py:8-259) -> mad_b200/structure_utils.py."""
from mad_b200.structure_utils import move_structure, move_copy_structure, refine_pdb, refine_poses, get_overlap  # noqa: F401
