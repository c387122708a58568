#ifndef REFSHIM_NUMAIF_H
#define REFSHIM_NUMAIF_H
#endif
