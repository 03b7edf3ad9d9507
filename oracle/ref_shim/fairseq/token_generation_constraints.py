class ConstraintState: pass
class OrderedConstraintState(ConstraintState): pass
class UnorderedConstraintState(ConstraintState): pass
