#include <petsc.h>
