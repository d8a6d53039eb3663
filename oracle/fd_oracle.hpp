/*
 * ORACLE — test infrastructure, NOT product code.
 *
 * Literal CPU restatement of alert GENERATION, one object per edge failure detector
 * exactly like the Java:
 *
 *   oracle::PingPongFailureDetector <- rapid/src/main/java/com/vrg/rapid/monitoring/impl/PingPongFailureDetector.java:38-121
 *   oracle::FdNode                  <- MembershipService.java:697-707 (createFailureDetectorsForCurrentConfiguration:
 *                                      one detector per entry of getSubjectsOf(myAddr), in ring order),
 *                                      :472-495 (edgeFailureNotification: the AlertMessage of a notification),
 *                                      :690-692 (createNotifierForSubject)
 *
 * The network is replaced by a probe-outcome callback (the reference's tests do the same with
 * StaticFailureDetector / blocked-server interceptors): the ProbeCallback of a probe sent in run() is
 * resolved before the next run() of the same detector.  Only tests/ may use this.
 */
#ifndef RAPID_FD_ORACLE_HPP
#define RAPID_FD_ORACLE_HPP

#include <functional>
#include <vector>

#include "rapid_oracle.hpp"

namespace oracle {

enum ProbeOutcome : int32_t { PROBE_OK = 0, PROBE_FAILED = 1, PROBE_BOOTSTRAPPING = 2 };

class PingPongFailureDetector {
public:
    static const int FAILURE_THRESHOLD = 10;                  /* :41 */
    static const int BOOTSTRAP_COUNT_THRESHOLD = 30;          /* :45 */

    explicit PingPongFailureDetector(const Endpoint& subject) : subject_(subject) {}

    /* :75-85 run() — returns true iff notifier.run() fired; otherwise one probe is sent and `probe` resolves it */
    bool run(const std::function<ProbeOutcome(const Endpoint&)>& probe) {
        if (hasFailed() && !notified_) {                      /* :76-79 */
            notified_ = true;
            return true;
        }
        const ProbeOutcome o = probe(subject_);               /* :80-84 sendMessageBestEffort + ProbeCallback */
        if (o == PROBE_FAILED) handleProbeOnFailure();        /* :110-112 onFailure */
        else if (o == PROBE_BOOTSTRAPPING) {                  /* :97-104 */
            const int n = ++bootstrapResponseCount_;
            if (n > BOOTSTRAP_COUNT_THRESHOLD) handleProbeOnFailure();
        }                                                     /* else handleProbeOnSuccess(): nothing (:115-117) */
        return false;
    }

    const Endpoint& subject() const { return subject_; }
    int failureCount() const { return failureCount_; }
    bool notified() const { return notified_; }

private:
    bool hasFailed() const { return failureCount_ >= FAILURE_THRESHOLD; }    /* :71-73 */
    void handleProbeOnFailure() { ++failureCount_; }                         /* :120-123 */

    Endpoint subject_;
    int failureCount_ = 0, bootstrapResponseCount_ = 0;
    bool notified_ = false;
};

/* One node's failure detectors for the current configuration and the alerts they raise. */
class FdNode {
public:
    FdNode(const MembershipView* view, const Endpoint& myAddr) : view_(view), myAddr_(myAddr) {
        for (const Endpoint& s : view->getSubjectsOf(myAddr)) fds_.emplace_back(s);      /* MembershipService.java:698-699 */
    }

    /* One failure-detector interval: every detector's run() in creation order.  Appends the AlertMessages the notifiers
     * enqueue (edgeFailureNotification :484-491: DOWN, all ring numbers of the edge, the current configuration id). */
    void tick(const std::function<ProbeOutcome(const Endpoint&, const Endpoint&)>& probe, int64_t configurationId,
              std::vector<AlertMessage>* out) {
        for (PingPongFailureDetector& fd : fds_) {
            const bool fired = fd.run([&](const Endpoint& s) { return probe(myAddr_, s); });
            if (!fired) continue;
            AlertMessage m;
            m.edgeSrc = myAddr_;
            m.edgeDst = fd.subject();
            m.edgeStatus = DOWN;
            m.configurationId = configurationId;
            for (int r : view_->getRingNumbers(myAddr_, fd.subject())) m.ringNumber.push_back(r);
            out->push_back(m);
        }
    }

    const std::vector<PingPongFailureDetector>& detectors() const { return fds_; }

private:
    const MembershipView* view_;
    Endpoint myAddr_;
    std::vector<PingPongFailureDetector> fds_;
};

}  // namespace oracle

#endif
