#include "engine.h"
struct S2melState {};
void s2mel_destroy(S2melState* s) { delete s; }
