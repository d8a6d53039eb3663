/*
 * ORACLE — test infrastructure, NOT product code.
 *
 * Literal CPU restatement of the reference's classic-Paxos fallback, one class
 * per node exactly like the Java:
 *
 *   oracle::ClassicPaxos  <-  rapid/src/main/java/com/vrg/rapid/Paxos.java (whole file)
 *   oracle::Rank          <-  rapid/src/main/proto/rapid.proto:133-137
 *   Phase1a/1b/2a/2b      <-  rapid.proto:139-169
 *
 * Only tests/ may use it.  Outgoing messages are RETURNED (the Java hands them to an IBroadcaster /
 * IMessagingClient); who they go to is noted per method.  `myAddr.hashCode()` (Paxos.java:102) is a
 * protobuf-generated hash that cannot be reproduced without the JVM, so the node index is a
 * constructor argument — only its ORDER between nodes matters to the protocol (compareRanks :333-339).
 *
 * Pinning: the reference's PaxosTests.java tables — coordinatorRuleTests (19 rows, :257-295),
 * coordinatorRuleTestsSameRank (17 rows, :362-393), the 8 mixed-value recovery rows (:176-192) and
 * the nValues cases (:128-136) — are ported in tests/test_oracle_classic_paxos.py.
 */
#ifndef RAPID_PAXOS_ORACLE_HPP
#define RAPID_PAXOS_ORACLE_HPP

#include <map>
#include <set>
#include <stdexcept>
#include <vector>

#include "rapid_oracle.hpp"

namespace oracle {

struct Rank {                                     /* rapid.proto:133-137 */
    int32_t round = 0, nodeIndex = 0;
    bool operator==(const Rank& o) const { return round == o.round && nodeIndex == o.nodeIndex; }
    bool operator<(const Rank& o) const {         // std::map key only (HashMap<Rank,..> in the Java)
        return round != o.round ? round < o.round : nodeIndex < o.nodeIndex;
    }
};

/* Paxos.java:333-339 — primary by round, secondary by node index (both signed int compares) */
inline int compareRanks(const Rank& left, const Rank& right) {
    if (left.round != right.round) return left.round < right.round ? -1 : 1;
    if (left.nodeIndex != right.nodeIndex) return left.nodeIndex < right.nodeIndex ? -1 : 1;
    return 0;
}

typedef std::vector<Endpoint> Value;

struct Phase1aMessage { Endpoint sender; int64_t configurationId = 0; Rank rank; };
struct Phase1bMessage { Endpoint sender; int64_t configurationId = 0; Rank rnd, vrnd; Value vval; };
struct Phase2aMessage { Endpoint sender; int64_t configurationId = 0; Rank rnd; Value vval; };
struct Phase2bMessage { Endpoint sender; int64_t configurationId = 0; Rank rnd; Value endpoints; };

class ClassicPaxos {
public:
    /* Paxos.java:76-90 */
    ClassicPaxos(const Endpoint& myAddr, int32_t myAddrHashCode, int64_t configurationId, int N)
        : configurationId_(configurationId), myAddr_(myAddr), myHash_(myAddrHashCode), N_(N) {}

    /* :98-113 startPhase1a — true iff a Phase1aMessage is broadcast to every member */
    bool startPhase1a(int round, Phase1aMessage* out) {
        if (crnd_.round > round) return false;                               /* :99-101 */
        crnd_.round = round;                                                 /* :102 */
        crnd_.nodeIndex = myHash_;
        out->sender = myAddr_; out->configurationId = configurationId_; out->rank = crnd_;
        return true;
    }

    /* :120-151 handlePhase1aMessage — true iff a Phase1bMessage is sent back to m.sender */
    bool handlePhase1aMessage(const Phase1aMessage& m, Phase1bMessage* reply) {
        if (m.configurationId != configurationId_) return false;             /* :121-123 */
        if (compareRanks(rnd_, m.rank) < 0) rnd_ = m.rank;                   /* :125-127 */
        else return false;                                                   /* :128-134 */
        reply->sender = myAddr_; reply->configurationId = configurationId_;  /* :138-144 */
        reply->rnd = rnd_; reply->vrnd = vrnd_; reply->vval = vval_;
        return true;
    }

    /* :159-191 handlePhase1bMessage — true iff a Phase2aMessage is broadcast */
    bool handlePhase1bMessage(const Phase1bMessage& m, Phase2aMessage* out) {
        if (m.configurationId != configurationId_) return false;             /* :160-162 */
        if (compareRanks(crnd_, m.rnd) != 0) return false;                   /* :165-167 */
        phase1bMessages_.push_back(m);                                       /* :171 */
        if ((int)phase1bMessages_.size() > (N_ / 2)) {                       /* :173 */
            const Value chosen = selectProposalUsingCoordinatorRule(phase1bMessages_);
            if (crnd_ == m.rnd && cval_.empty() && !chosen.empty()) {        /* :177 */
                cval_ = chosen;
                out->sender = myAddr_; out->configurationId = configurationId_;
                out->rnd = crnd_; out->vval = chosen;
                return true;
            }
        }
        return false;
    }

    /* :198-216 handlePhase2aMessage — true iff a Phase2bMessage is broadcast */
    bool handlePhase2aMessage(const Phase2aMessage& m, Phase2bMessage* out) {
        if (m.configurationId != configurationId_) return false;             /* :199-201 */
        if (compareRanks(rnd_, m.rnd) <= 0 && !(vrnd_ == m.rnd)) {           /* :204 */
            rnd_ = m.rnd; vrnd_ = m.rnd; vval_ = m.vval;
            out->sender = myAddr_; out->configurationId = configurationId_;
            out->rnd = m.rnd; out->endpoints = vval_;
            return true;
        }
        return false;
    }

    /* :223-236 handlePhase2bMessage — true iff THIS message made the node decide */
    bool handlePhase2bMessage(const Phase2bMessage& m) {
        if (m.configurationId != configurationId_) return false;             /* :224-226 */
        std::map<Endpoint, Value>& inRnd = acceptResponses_[m.rnd];          /* :228-229 */
        inRnd[m.sender] = m.endpoints;                                       /* :230 put (overwrites) */
        if ((int)inRnd.size() > (N_ / 2) && !decided_) {                     /* :231 */
            decision_ = m.endpoints;                                         /* :232 the ARRIVING message's list */
            decided_ = true;
            return true;
        }
        return false;
    }

    /* :244-257 registerFastRoundVote */
    void registerFastRoundVote(const Value& vote) {
        if (rnd_.round > 1) return;                                          /* :246-248 */
        rnd_.round = 1; rnd_.nodeIndex = 1;                                  /* :254 */
        vrnd_ = rnd_;
        vval_ = vote;
    }

    /* :271-328 selectProposalUsingCoordinatorRule */
    Value selectProposalUsingCoordinatorRule(const std::vector<Phase1bMessage>& msgs) const {
        if (msgs.empty()) throw std::invalid_argument("phase1bMessages was empty");   /* :274 */
        Rank maxVrndSoFar = msgs[0].vrnd;                                    /* :272-274 */
        for (const Phase1bMessage& m : msgs)
            if (compareRanks(m.vrnd, maxVrndSoFar) > 0) maxVrndSoFar = m.vrnd;
        std::vector<const Value*> collectedVvals;                            /* :278-282 */
        for (const Phase1bMessage& m : msgs)
            if (m.vrnd == maxVrndSoFar && !m.vval.empty()) collectedVvals.push_back(&m.vval);
        std::set<Value> setOfCollectedVvals;                                 /* :283 */
        for (const Value* v : collectedVvals) setOfCollectedVvals.insert(*v);
        const Value* chosen = nullptr;
        if (setOfCollectedVvals.size() == 1) {                               /* :287-289 */
            chosen = collectedVvals.front();
        } else if (collectedVvals.size() > 1) {                              /* :293-308 */
            std::map<Value, int> counters;
            for (const Value* v : collectedVvals) {
                int& count = counters[*v];                                   // absent => 0
                if (count + 1 > (N_ / 4)) { chosen = v; break; }
                count = count + 1;
            }
        }
        if (chosen == nullptr) {                                             /* :318-326 first non-empty vval */
            for (const Phase1bMessage& m : msgs)
                if (!m.vval.empty()) return m.vval;
            return Value();
        }
        return *chosen;
    }

    bool decided() const { return decided_; }
    const Value& decision() const { return decision_; }
    Rank rnd() const { return rnd_; }
    Rank vrnd() const { return vrnd_; }
    Rank crnd() const { return crnd_; }
    const Value& vval() const { return vval_; }
    const Value& cval() const { return cval_; }
    size_t numPhase1bMessages() const { return phase1bMessages_.size(); }

private:
    int64_t configurationId_;
    Endpoint myAddr_;
    int32_t myHash_;
    int N_;
    Rank rnd_, vrnd_;
    Value vval_;
    std::vector<Phase1bMessage> phase1bMessages_;
    std::map<Rank, std::map<Endpoint, Value>> acceptResponses_;
    Rank crnd_;
    Value cval_;
    bool decided_ = false;
    Value decision_;
};

}  // namespace oracle

#endif
