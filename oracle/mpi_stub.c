/* mpi_stub.c -- TEST INFRASTRUCTURE (oracle).  Not part of the product.
 * The in-process MPI of oracle/mpi_stub/mpi.h: ranks are pthreads of one process. */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "mpi.h"

#define MAX_RANKS 64
#define MAX_COMMS 16

typedef struct message {
    int src, tag, comm;
    size_t bytes;
    void *data;
    struct message *next;
} message;

struct mpi_stub_request {
    int is_recv, source, tag, comm;
    void *buf;
    size_t bytes;
};

static int g_nranks = 1;
static __thread int t_rank = 0;
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_arrived = PTHREAD_COND_INITIALIZER;
static message *g_inbox_head[MAX_RANKS], *g_inbox_tail[MAX_RANKS];   /* per destination, in send order */

/* communicators: 0 = world; Cartesian ones remember their dimensions */
static struct { int ndims, dims[3]; } g_comm[MAX_COMMS];
static int g_ncomm = 1;

/* a reusable barrier + a slot per rank for the reductions */
static pthread_mutex_t g_bar_lock = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_bar_cond = PTHREAD_COND_INITIALIZER;
static int g_bar_count = 0;
static unsigned long g_bar_generation = 0;
static double g_slot[MAX_RANKS][4];

void mpi_stub_world(int nranks)
{
    if (nranks < 1 || nranks > MAX_RANKS) { fprintf(stderr, "mpi_stub: %d ranks\n", nranks); abort(); }
    g_nranks = nranks;
    g_ncomm = 1;
}
void mpi_stub_set_rank(int rank) { t_rank = rank; }

typedef struct { int rank; void (*fn)(int, void *); void *arg; } launch;
static void *trampoline(void *p)
{
    launch *l = (launch *)p;
    t_rank = l->rank;
    l->fn(l->rank, l->arg);
    return NULL;
}
int mpi_stub_run(int nranks, void (*fn)(int rank, void *arg), void *arg)
{
    pthread_t th[MAX_RANKS];
    launch l[MAX_RANKS];
    mpi_stub_world(nranks);
    for (int r = 0; r < nranks; r++) {
        l[r].rank = r; l[r].fn = fn; l[r].arg = arg;
        if (pthread_create(&th[r], NULL, trampoline, &l[r])) return -1;
    }
    for (int r = 0; r < nranks; r++) pthread_join(th[r], NULL);
    return 0;
}

int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int *size) { (void)comm; *size = g_nranks; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { (void)comm; *rank = t_rank; return MPI_SUCCESS; }
double MPI_Wtime(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

/* collective: every rank calls it with the same arguments and gets the same communicator id */
int MPI_Cart_create(MPI_Comm comm, int ndims, const int *dims, const int *periods, int reorder, MPI_Comm *cart)
{
    (void)comm; (void)reorder;
    int cells = 1;
    for (int d = 0; d < ndims; d++) {
        cells *= dims[d];
        if (periods[d]) { fprintf(stderr, "mpi_stub: periodic grids are not modelled\n"); abort(); }
    }
    if (ndims > 3 || cells != g_nranks) {
        fprintf(stderr, "mpi_stub: a %d-cell Cartesian grid over %d ranks (MPI would fail here too)\n", cells, g_nranks);
        abort();
    }
    MPI_Barrier(comm);
    pthread_mutex_lock(&g_lock);
    if (t_rank == 0) {
        if (g_ncomm >= MAX_COMMS) g_ncomm = 1;
        g_comm[g_ncomm].ndims = ndims;
        for (int d = 0; d < 3; d++) g_comm[g_ncomm].dims[d] = d < ndims ? dims[d] : 1;
        g_ncomm++;
    }
    pthread_mutex_unlock(&g_lock);
    MPI_Barrier(comm);
    *cart = g_ncomm - 1;
    return MPI_SUCCESS;
}

/* row-major ranks (last dimension fastest), non-periodic: off the grid is MPI_PROC_NULL */
int MPI_Cart_shift(MPI_Comm comm, int direction, int disp, int *rank_source, int *rank_dest)
{
    const int *dims = g_comm[comm].dims;
    int c[3] = { t_rank / (dims[1] * dims[2]), (t_rank / dims[2]) % dims[1], t_rank % dims[2] };
    for (int side = 0; side < 2; side++) {
        int at[3] = { c[0], c[1], c[2] };
        at[direction] += side == 0 ? -disp : disp;
        int r = (at[direction] < 0 || at[direction] >= dims[direction]) ? MPI_PROC_NULL
                                                                        : (at[0] * dims[1] + at[1]) * dims[2] + at[2];
        if (side == 0) *rank_source = r; else *rank_dest = r;
    }
    return MPI_SUCCESS;
}

int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype *newtype)
{
    *newtype = (0x7f << 24) | (count * (oldtype & 0xffffff));
    return MPI_SUCCESS;
}
int MPI_Type_commit(MPI_Datatype *type) { (void)type; return MPI_SUCCESS; }

int MPI_Isend(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request *req)
{
    struct mpi_stub_request *r = (struct mpi_stub_request *)calloc(1, sizeof *r);
    *req = r;
    if (dest == MPI_PROC_NULL) return MPI_SUCCESS;
    message *m = (message *)malloc(sizeof *m);
    m->src = t_rank; m->tag = tag; m->comm = comm; m->next = NULL;
    m->bytes = (size_t)count * (size_t)(type & 0xffffff);
    m->data = malloc(m->bytes ? m->bytes : 1);
    memcpy(m->data, buf, m->bytes);            /* eager: the message leaves now */
    pthread_mutex_lock(&g_lock);
    if (g_inbox_tail[dest]) g_inbox_tail[dest]->next = m; else g_inbox_head[dest] = m;
    g_inbox_tail[dest] = m;
    pthread_cond_broadcast(&g_arrived);
    pthread_mutex_unlock(&g_lock);
    return MPI_SUCCESS;
}

int MPI_Irecv(void *buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request *req)
{
    struct mpi_stub_request *r = (struct mpi_stub_request *)calloc(1, sizeof *r);
    r->is_recv = source != MPI_PROC_NULL;
    r->source = source; r->tag = tag; r->comm = comm; r->buf = buf;
    r->bytes = (size_t)count * (size_t)(type & 0xffffff);
    *req = r;
    return MPI_SUCCESS;
}

int MPI_Wait(MPI_Request *req, MPI_Status *status)
{
    struct mpi_stub_request *r = *req;
    if (r && r->is_recv) {
        pthread_mutex_lock(&g_lock);
        for (;;) {
            message *prev = NULL, *m = g_inbox_head[t_rank];
            while (m && !(m->src == r->source && m->tag == r->tag && m->comm == r->comm)) { prev = m; m = m->next; }
            if (m) {                              /* the OLDEST matching message: non-overtaking */
                if (prev) prev->next = m->next; else g_inbox_head[t_rank] = m->next;
                if (g_inbox_tail[t_rank] == m) g_inbox_tail[t_rank] = prev;
                pthread_mutex_unlock(&g_lock);
                if (m->bytes != r->bytes) { fprintf(stderr, "mpi_stub: message of %zu bytes for a receive of %zu\n", m->bytes, r->bytes); abort(); }
                memcpy(r->buf, m->data, m->bytes);
                if (status) { status->MPI_SOURCE = m->src; status->MPI_TAG = m->tag; status->MPI_ERROR = 0; }
                free(m->data); free(m);
                break;
            }
            pthread_cond_wait(&g_arrived, &g_lock);
        }
    }
    free(r);
    *req = NULL;
    return MPI_SUCCESS;
}

int MPI_Barrier(MPI_Comm comm)
{
    (void)comm;
    pthread_mutex_lock(&g_bar_lock);
    const unsigned long gen = g_bar_generation;
    if (++g_bar_count == g_nranks) {
        g_bar_count = 0;
        g_bar_generation++;
        pthread_cond_broadcast(&g_bar_cond);
    } else {
        while (gen == g_bar_generation) pthread_cond_wait(&g_bar_cond, &g_bar_lock);
    }
    pthread_mutex_unlock(&g_bar_lock);
    return MPI_SUCCESS;
}

/* MPI_SUM of float / long / int / double vectors of up to 4 elements, contributions added in rank order */
static void reduce_sum(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, int root_or_all)
{
    if (count > 4) { fprintf(stderr, "mpi_stub: reduction of %d elements\n", count); abort(); }
    for (int e = 0; e < count; e++) {
        switch (type) {
        case MPI_FLOAT: g_slot[t_rank][e] = ((const float *)sendbuf)[e]; break;
        case MPI_DOUBLE: g_slot[t_rank][e] = ((const double *)sendbuf)[e]; break;
        case MPI_LONG: g_slot[t_rank][e] = (double)((const long *)sendbuf)[e]; break;
        case MPI_INT: g_slot[t_rank][e] = ((const int *)sendbuf)[e]; break;
        default: fprintf(stderr, "mpi_stub: reduction of datatype %x\n", type); abort();
        }
    }
    MPI_Barrier(0);
    if (root_or_all < 0 || root_or_all == t_rank) {
        for (int e = 0; e < count; e++) {
            if (type == MPI_FLOAT) {
                float s = 0;   /* accumulated in the element type, rank 0 first */
                for (int r = 0; r < g_nranks; r++) s += (float)g_slot[r][e];
                ((float *)recvbuf)[e] = s;
            } else if (type == MPI_DOUBLE) {
                double s = 0;
                for (int r = 0; r < g_nranks; r++) s += g_slot[r][e];
                ((double *)recvbuf)[e] = s;
            } else if (type == MPI_LONG) {
                long s = 0;
                for (int r = 0; r < g_nranks; r++) s += (long)g_slot[r][e];
                ((long *)recvbuf)[e] = s;
            } else {
                int s = 0;
                for (int r = 0; r < g_nranks; r++) s += (int)g_slot[r][e];
                ((int *)recvbuf)[e] = s;
            }
        }
    }
    MPI_Barrier(0);
}

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm)
{
    (void)op; (void)comm;
    reduce_sum(sendbuf, recvbuf, count, type, -1);
    return MPI_SUCCESS;
}

int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm)
{
    (void)op; (void)comm;
    reduce_sum(sendbuf, recvbuf, count, type, root);
    return MPI_SUCCESS;
}
