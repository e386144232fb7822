// hostpipe.cuh -- the host-buffer step pipeline shared by mrmd_b200_md_run_host and mrmd_b200_slab_run_host.
//
// Per step: pos and vel come from (pinned) host buffers, one step runs, pos, vel and the scalars go back.  PCIe is full
// duplex and the positions are final before the force kernel starts, so the copies run on two extra streams and in
// chunks: the download of the positions overlaps the force kernel, and chunk c of the next step's upload starts as soon
// as chunk c of this step's download has landed, so both directions of the link stay busy.  Every byte of a host buffer
// is read only after the previous step's write to it.
#pragma once

#include <algorithm>
#include <cstdlib>

#include "handles.cuh"

namespace mrmd_b200
{
struct HostPipe
{
    static constexpr int MAX_CHUNKS = 8;
    cudaStream_t sIn = nullptr, sOut = nullptr;
    cudaEvent_t evUpPos = nullptr, evUpVel = nullptr, evPosReady = nullptr, evStepDone = nullptr;
    cudaEvent_t evDownPos[MAX_CHUNKS] = {}, evDownVel[MAX_CHUNKS] = {};
    DevBuf posIn, velIn, posOut, velOut;

    int init()
    {
        if (sIn != nullptr) return 0;
        // highest priority: the pack kernel in front of a download must not queue behind the blocks of the force
        // kernel it is meant to overlap
        int prioLow = 0, prioHigh = 0;
        MB_CUDA(cudaDeviceGetStreamPriorityRange(&prioLow, &prioHigh));
        MB_CUDA(cudaStreamCreateWithPriority(&sIn, cudaStreamNonBlocking, prioHigh));
        MB_CUDA(cudaStreamCreateWithPriority(&sOut, cudaStreamNonBlocking, prioHigh));
        for (cudaEvent_t* e : {&evUpPos, &evUpVel, &evPosReady, &evStepDone})
            MB_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        for (int c = 0; c < MAX_CHUNKS; ++c)
            for (cudaEvent_t* e : {&evDownPos[c], &evDownVel[c]}) MB_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        return 0;
    }
    void destroy()
    {
        for (cudaEvent_t e : {evUpPos, evUpVel, evPosReady, evStepDone})
            if (e != nullptr) cudaEventDestroy(e);
        for (int c = 0; c < MAX_CHUNKS; ++c)
            for (cudaEvent_t e : {evDownPos[c], evDownVel[c]})
                if (e != nullptr) cudaEventDestroy(e);
        if (sIn != nullptr) cudaStreamDestroy(sIn);
        if (sOut != nullptr) cudaStreamDestroy(sOut);
        sIn = sOut = nullptr;
        for (DevBuf* b : {&posIn, &velIn, &posOut, &velOut}) b->release();
    }
};

// stepFn(posReadyEvent): one step on `st` that records the event once the positions (and the atom order) are final and
// ends with the velocities final (no deferred postForceIntegrate); the number of local atoms may change in a step
// (slab migration), the host buffers must hold the largest count.
template <class StepFn>
int hostPipeRun(HostPipe& hp, mrmd_b200_atoms* a, int64_t nsteps, double* posHost, double* velHost, double* scalarsHost,
                const double* dRes, const double* maxDisplacement, bool writeCount, cudaStream_t st, StepFn stepFn)
{
    MB_TRY(hp.init());
    // chunks per array and direction (MRMD_B200_HB_CHUNKS overrides the default of 4 for measurements)
    int K = 4;
    if (const char* e = std::getenv("MRMD_B200_HB_CHUNKS")) K = std::max(1, std::min(int(HostPipe::MAX_CHUNKS), std::atoi(e)));
    size_t off[HostPipe::MAX_CHUNKS + 1];
    auto setChunks = [&](int64_t n)
    {
        const size_t bytes = size_t(n) * 24;
        for (int c = 0; c <= K; ++c) off[c] = (c == K) ? bytes : ((bytes * size_t(c) / K) & ~size_t(255));
    };
    // one direction of one array: chunk c waits for gate[c] (if any) and records done[c] (always, so that a later
    // wait on it never sees a stale event)
    auto copyChunks = [&](void* dst, const void* src, cudaMemcpyKind kind, cudaStream_t cs, cudaEvent_t* gate,
                          cudaEvent_t* done) -> int
    {
        for (int c = 0; c < K; ++c)
        {
            if (gate != nullptr) MB_CUDA(cudaStreamWaitEvent(cs, gate[c], 0));
            if (off[c + 1] > off[c])
                MB_CUDA(cudaMemcpyAsync(static_cast<char*>(dst) + off[c], static_cast<const char*>(src) + off[c],
                                        off[c + 1] - off[c], kind, cs));
            if (done != nullptr) MB_CUDA(cudaEventRecord(done[c], cs));
        }
        return 0;
    };
    int rc = 0;
    auto step = [&](int64_t i) -> int
    {
        const int64_t nIn = a->numLocal;
        setChunks(nIn);
        for (DevBuf* b : {&hp.posIn, &hp.velIn}) MB_TRY(b->reserve(std::max<size_t>(size_t(nIn) * 24, 8)));
        // host -> device: this step's inputs (the host buffers were last written by the previous step's download)
        MB_TRY(copyChunks(hp.posIn.p, posHost, cudaMemcpyHostToDevice, hp.sIn, i > 0 ? hp.evDownPos : nullptr, nullptr));
        MB_CUDA(cudaEventRecord(hp.evUpPos, hp.sIn));
        MB_TRY(copyChunks(hp.velIn.p, velHost, cudaMemcpyHostToDevice, hp.sIn, i > 0 ? hp.evDownVel : nullptr, nullptr));
        MB_CUDA(cudaEventRecord(hp.evUpVel, hp.sIn));
        MB_CUDA(cudaStreamWaitEvent(st, hp.evUpPos, 0));
        MB_TRY(atomsFieldFromDense(a, MRMD_B200_ATOM_POS, hp.posIn.as<double>(), nIn, st));
        MB_CUDA(cudaStreamWaitEvent(st, hp.evUpVel, 0));
        MB_TRY(atomsFieldFromDense(a, MRMD_B200_ATOM_VEL, hp.velIn.as<double>(), nIn, st));
        MB_TRY(stepFn(hp.evPosReady));  // records evPosReady in front of the force kernel
        const int64_t nOut = a->numLocal;
        setChunks(nOut);
        for (DevBuf* b : {&hp.posOut, &hp.velOut}) MB_TRY(b->reserve(std::max<size_t>(size_t(nOut) * 24, 8)));
        // device -> host: positions while the force kernel runs ...
        MB_CUDA(cudaStreamWaitEvent(hp.sOut, hp.evPosReady, 0));
        MB_TRY(atomsFieldToDense(a, MRMD_B200_ATOM_POS, hp.posOut.as<double>(), nOut, hp.sOut));
        MB_TRY(copyChunks(posHost, hp.posOut.p, cudaMemcpyDeviceToHost, hp.sOut, nullptr, hp.evDownPos));
        // ... velocities and scalars after postForceIntegrate
        MB_CUDA(cudaEventRecord(hp.evStepDone, st));
        MB_CUDA(cudaStreamWaitEvent(hp.sOut, hp.evStepDone, 0));
        MB_TRY(atomsFieldToDense(a, MRMD_B200_ATOM_VEL, hp.velOut.as<double>(), nOut, hp.sOut));
        MB_TRY(copyChunks(velHost, hp.velOut.p, cudaMemcpyDeviceToHost, hp.sOut, nullptr, hp.evDownVel));
        if (scalarsHost != nullptr)
        {
            MB_CUDA(cudaMemcpyAsync(scalarsHost, dRes, 16, cudaMemcpyDeviceToHost, hp.sOut));
            scalarsHost[2] = *maxDisplacement;
            if (writeCount) scalarsHost[3] = static_cast<double>(nOut);
        }
        return 0;
    };
    for (int64_t i = 0; i < nsteps && rc == 0; ++i) rc = step(i);
    // the caller's stream must not run ahead of the copies it ordered (the next call reuses the staging buffers)
    const cudaError_t e1 = cudaStreamSynchronize(hp.sOut), e2 = cudaStreamSynchronize(hp.sIn);
    if (rc != 0) return rc;
    MB_CUDA(e1);
    MB_CUDA(e2);
    return 0;
}
}  // namespace mrmd_b200
