/*
 * Drop-in physical operator: replaces the INSIDE of the partial-aggregation stage
 *
 *     SnappyHashAggregateExec(Partial) <- [ProjectExec] <- [FilterExec] <- ColumnTableScan
 *
 * by one fused GPU execution per partition that emits the very same UnsafeRow(groupKeys ++ aggBuffers)
 * rows, so the Exchange + SnappyHashAggregateExec(Final) / CollectAggregateExec above it are untouched
 * (SURVEY.md 8b; core/.../aggregate/SnappyHashAggregateExec.scala:1148-1178; CollectAggregateExec.scala:67-121).
 *
 * NOT COMPILED in this repository's container (no JDK/Scala); a sketch against the reference's class
 * names for a maintainer to adapt.  Registration (core/hive/SnappySessionState.scala:699-707, 730-739):
 *
 *     override def queryPreparations: Seq[Rule[SparkPlan]] = super.queryPreparations :+ GpuOffloadRule(session)
 *
 * The node keeps the children so EXPLAIN, plan caching (tokenised ParamLiterals are read per execution),
 * CollapseCollocatedPlans and the SQL UI see the original ColumnTableScan / SnappyHashAggregate names.
 */
package io.snappydata.gpu

import org.apache.spark.rdd.RDD
import org.apache.spark.sql.catalyst.InternalRow
import org.apache.spark.sql.catalyst.expressions._
import org.apache.spark.sql.catalyst.expressions.aggregate._
import org.apache.spark.sql.catalyst.rules.Rule
import org.apache.spark.sql.execution._
import org.apache.spark.sql.execution.aggregate.SnappyHashAggregateExec
import org.apache.spark.sql.execution.columnar.{ColumnBatchIterator, ColumnTableScan}
import org.apache.spark.sql.execution.metric.SQLMetrics
import org.apache.spark.unsafe.Platform

/** Matches the supported shape and swaps it; anything else is left to the stock operators (that is the
  * planner deciding, not a run-time CPU fallback: once planned, failures surface as exceptions). */
case class GpuOffloadRule(session: org.apache.spark.sql.SparkSession) extends Rule[SparkPlan] {
  override def apply(plan: SparkPlan): SparkPlan = plan.transformUp {
    case agg: SnappyHashAggregateExec
      if agg.aggregateExpressions.forall(_.mode == Partial) && SnappyGpuNative.isLoaded =>
      GpuPlanSerializer.tryBuild(agg) match {
        case Some(desc) => GpuPartialAggregateExec(agg, desc)
        case None => agg
      }
  }
}

case class GpuPartialAggregateExec(original: SnappyHashAggregateExec, desc: GpuPlanDesc)
    extends UnaryExecNode {

  override def child: SparkPlan = original.child
  override def output: Seq[Attribute] = original.output
  override def nodeName: String = original.nodeName // "SnappyHashAggregate" (SnappyHashAggregateExec.scala:110-111)

  // same SQLMetrics as the two operators it fuses (ColumnTableScan.scala:111-127, SnappyHashAggregateExec.scala:132-137)
  override lazy val metrics = original.metrics ++ Map(
    "columnBatchesSeen" -> SQLMetrics.createMetric(sparkContext, "column batches seen"),
    "columnBatchesSkipped" -> SQLMetrics.createMetric(sparkContext, "column batches skipped by the predicate"),
    "numRowsBuffer" -> SQLMetrics.createMetric(sparkContext, "number of output rows from row buffer"))

  override protected def doExecute(): RDD[InternalRow] = {
    val scan = desc.scan // the ColumnTableScan under the filter/project
    // each partition iterator yields exactly two elements: the row-buffer iterator, then the
    // ColumnBatchIterator (ColumnTableScan.scala:236-241)
    scan.dataRDD.mapPartitionsWithIndex { (partition, iter) =>
      SnappyGpuNative.init(partition % GpuPlanSerializer.numDevices)
      val plan = SnappyGpuNative.planCreate(desc.address)
      try {
        SnappyGpuNative.planSetLiterals(plan, desc.literalsAddress(), desc.numLiterals) // ParamLiteral values of THIS execution
        val rowBuffer = iter.next().asInstanceOf[Iterator[InternalRow]]
        GpuPlanSerializer.submitRowBuffer(plan, rowBuffer, desc)
        val batches = iter.next().asInstanceOf[ColumnBatchIterator]
        while (batches.hasNext) {
          if (org.apache.spark.TaskContext.get().isInterrupted()) throw new org.apache.spark.TaskKilledException
          val stats = batches.next() // stats row buffer (ColumnBatchIterator.scala:179-223)
          GpuPlanSerializer.submitBatch(plan, batches, stats, desc) // getColumnLob / delta / delete buffers -> batchSubmit
        }
        GpuPlanSerializer.finishToUnsafeRows(plan, desc, longMetric("numOutputRows"))
      } finally SnappyGpuNative.planDestroy(plan)
    }
  }
}
