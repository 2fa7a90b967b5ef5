/*
 * Drop-in physical operator: replaces the INSIDE of the partial-aggregation stage
 *
 *     SnappyHashAggregateExec(Partial) <- [ProjectExec] <- [FilterExec] <- ColumnTableScan
 *
 * by one fused GPU execution per partition that emits the very same UnsafeRow(groupKeys ++ aggBuffers) rows, so the
 * Exchange + SnappyHashAggregateExec(Final) / CollectAggregateExec above it are untouched (SURVEY.md 8b;
 * core/.../aggregate/SnappyHashAggregateExec.scala:1148-1178; CollectAggregateExec.scala:67-121).
 *
 * NOT COMPILED in this repository's container (no JDK / scalac); written against the reference's signatures.
 * Registration (core/hive/SnappySessionState.scala:699-707, 730-739):
 *
 *     override def queryPreparations: Seq[Rule[SparkPlan]] = super.queryPreparations :+ GpuOffloadRule(session)
 *
 * (before CollapseCodegenStages, so the replaced subtree is not wrapped into a WholeStageCodegenExec).  The node keeps the
 * original aggregate so EXPLAIN, plan caching (ParamLiteral values are read per execution), CollapseCollocatedPlans and
 * the SQL UI see the reference's node names.
 */
package org.apache.spark.sql.execution.columnar.gpu

import io.snappydata.gpu.SnappyGpuNative

import org.apache.spark.{TaskContext, TaskKilledException}
import org.apache.spark.rdd.RDD
import org.apache.spark.sql.SparkSession
import org.apache.spark.sql.catalyst.InternalRow
import org.apache.spark.sql.catalyst.expressions.Attribute
import org.apache.spark.sql.catalyst.expressions.aggregate.Partial
import org.apache.spark.sql.catalyst.rules.Rule
import org.apache.spark.sql.execution.{SparkPlan, UnaryExecNode}
import org.apache.spark.sql.execution.aggregate.SnappyHashAggregateExec
import org.apache.spark.sql.execution.columnar.ColumnBatchIterator
import org.apache.spark.sql.execution.metric.SQLMetrics

/** Matches the supported shape and swaps it; anything else is left to the stock operators (the planner deciding, not a
  * run-time CPU fallback: once planned, failures surface as exceptions). */
case class GpuOffloadRule(session: SparkSession) extends Rule[SparkPlan] {
  override def apply(plan: SparkPlan): SparkPlan = if (!SnappyGpuNative.isLoaded) plan else plan.transformUp {
    case agg: SnappyHashAggregateExec if agg.aggregateExpressions.nonEmpty &&
        agg.aggregateExpressions.forall(_.mode == Partial) =>
      GpuPlanSerializer.tryBuild(agg) match {
        case Some(desc) => GpuPartialAggregateExec(agg, desc)
        case None => agg
      }
  }
}

case class GpuPartialAggregateExec(original: SnappyHashAggregateExec, @transient desc: GpuPlanDesc)
    extends UnaryExecNode {

  override def child: SparkPlan = original.child
  override def output: Seq[Attribute] = original.output
  override def nodeName: String = original.nodeName // "SnappyHashAggregate" | "BufferMapHashAggregate" (SnappyHashAggregateExec.scala:110-111)

  // the SQLMetrics of the two operators it fuses (ColumnTableScan.scala:111-127, SnappyHashAggregateExec.scala:132-137)
  override lazy val metrics = original.metrics ++ Map(
    "numRowsBuffer" -> SQLMetrics.createMetric(sparkContext, "number of output rows from row buffer"),
    "columnBatchesSeen" -> SQLMetrics.createMetric(sparkContext, "column batches seen"),
    "updatedColumnCount" -> SQLMetrics.createMetric(sparkContext, "total updated columns in batches"),
    "deletedBatchCount" -> SQLMetrics.createMetric(sparkContext, "column batches having deletes"),
    "columnBatchesSkipped" -> SQLMetrics.createMetric(sparkContext, "column batches skipped by the predicate"))

  override protected def doExecute(): RDD[InternalRow] = {
    val planDesc = desc
    val numOutputRows = longMetric("numOutputRows")
    val metricNames = Array("numRowsBuffer", "columnBatchesSeen", "updatedColumnCount", "deletedBatchCount", "columnBatchesSkipped")
    val ms = metricNames.map(longMetric)
    val aggTime = longMetric("aggTime")
    // each partition iterator yields exactly two elements: the row-buffer iterator (a ResultSetTraversal), then the
    // ColumnBatchIterator (ColumnTableScan.scala:236-241)
    planDesc.scan.dataRDD.mapPartitionsWithIndex { (partition, iter) =>
      SnappyGpuNative.init(partition % GpuPlanSerializer.numDevices)     // one GPU per partition, round robin
      val plan = SnappyGpuNative.planCreate(planDesc.address)
      try {
        val lits = planDesc.writeLiterals()                               // ParamLiteral values of THIS execution
        try SnappyGpuNative.planSetLiterals(plan, lits.address, planDesc.numLiterals) finally lits.free()
        GpuPlanSerializer.submitRowBuffer(plan, iter.next().asInstanceOf[Iterator[_]], planDesc)
        val batches = iter.next().asInstanceOf[ColumnBatchIterator]
        val ctx = TaskContext.get()
        while (batches.hasNext) {
          if ((ctx ne null) && ctx.isInterrupted()) throw new TaskKilledException   // ColumnBatch.scala:63-66
          val stats = batches.next()                                      // stats row of the batch (ColumnBatchIterator.scala:179-223)
          GpuPlanSerializer.submitBatch(plan, batches, stats, planDesc)
        }
        val rows = GpuPlanSerializer.finishToUnsafeRows(plan, planDesc, numOutputRows)
        val m = new Array[Long](12)
        SnappyGpuNative.planMetrics(plan, m)                              // include/snappy_gpu.h: sd_plan_metrics
        ms(0).add(m(1)); ms(1).add(m(2)); ms(2).add(m(3)); ms(3).add(m(4)); ms(4).add(m(5))
        aggTime.add(m(6) / 1000000L)
        rows
      } finally SnappyGpuNative.planDestroy(plan)
    }
  }
}
