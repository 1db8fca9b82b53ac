"""Bitwise ufuncs (reference: cupy/_core/_routines_binary.pyx:4-133): integer and boolean loops only, the
and / or / xor ones carrying the scatter operation `ufunc.at` applies."""
from cupy_b200._core._kernel import create_ufunc

_INT2 = ('bb->b', 'BB->B', 'hh->h', 'HH->H', 'ii->i', 'II->I', 'll->l', 'LL->L', 'qq->q', 'QQ->Q')
_INT1 = ('b->b', 'B->B', 'h->h', 'H->H', 'i->i', 'I->I', 'l->l', 'L->L', 'q->q', 'Q->Q')


def _create_bit_op(name, op, no_bool, doc='', scatter_op=None):
    return create_ufunc('cupy_' + name, (() if no_bool else ('??->?',)) + _INT2, 'out0 = in0 %s in1' % op,
                        doc=doc, scatter_op=scatter_op)


bitwise_and = _create_bit_op('bitwise_and', '&', False, 'Computes the bitwise AND of two arrays elementwise.', 'and')
bitwise_or = _create_bit_op('bitwise_or', '|', False, 'Computes the bitwise OR of two arrays elementwise.', 'or')
bitwise_xor = _create_bit_op('bitwise_xor', '^', False, 'Computes the bitwise XOR of two arrays elementwise.', 'xor')
invert = create_ufunc('cupy_invert', (('?->?', 'out0 = !in0'),) + _INT1, 'out0 = ~in0',
                      doc='Computes the bitwise NOT of an array elementwise.')
bitwise_not = invert
left_shift = _create_bit_op('left_shift', '<<', True, 'Shifts the bits of each integer element to the left.')
right_shift = _create_bit_op('right_shift', '>>', True, 'Shifts the bits of each integer element to the right.')
