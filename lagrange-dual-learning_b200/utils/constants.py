"""Link lengths of the planar arms, under the attribute names the reference's code reads
(reference: utils/constants.py -- every link of every arm is one unit long)."""

UNIT_LINK = 1.0


class Constants:
  """Manipulator geometry."""
  R2_LENGTH1 = R2_LENGTH2 = UNIT_LINK                    # two-link arm
  R3_LENGTH1 = R3_LENGTH2 = R3_LENGTH3 = UNIT_LINK       # three-link arm
  R8_LENGTH = UNIT_LINK                                  # eight-link arm (its FK never ran in the reference)
