from torchflows_b200.bijections.finite.matrix.permutation import (PermutationMatrix, RandomPermutationMatrix,
                                                                  ReversePermutationMatrix)
