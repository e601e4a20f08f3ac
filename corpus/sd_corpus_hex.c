/* MC / MN corpus generator — filled in with the MC/MN milestone. */
#include "../include/sdb200.h"
